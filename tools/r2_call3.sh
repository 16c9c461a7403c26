set -x
timeout 600 python -m pytest tests/test_gpu_packplan.py tests/test_gpu_gradcheck.py -q -x -s > gpurun_out/r2_c3_a.log 2>&1
tail -30 gpurun_out/r2_c3_a.log
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r2_c3_pytest.log 2>&1
tail -8 gpurun_out/r2_c3_pytest.log
python bench.py --quick --steps 5 --warmup 3 2>&1 | tail -1
FDG_WU_FAST=1 python bench.py --quick --steps 5 --warmup 3 2>&1 | tail -1
python tests/bench_configs.py 2>&1 | tail -5
