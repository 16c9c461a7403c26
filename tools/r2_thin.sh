set -x
python -m pytest tests/test_gpu_ops.py -x -q -k "thin or colsum" 2>&1 | tail -4 | tee gpurun_out/r2_thin_pytest.log
python tests/bench_thin.py 2>&1 | tee gpurun_out/r2_thin_bench.log
