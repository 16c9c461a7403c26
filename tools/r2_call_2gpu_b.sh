set -x
timeout 600 python -m pytest tests/test_gpu_dataparallel.py -q -s > gpurun_out/r2_2gpu_dp_pytest.log 2>&1
tail -5 gpurun_out/r2_2gpu_dp_pytest.log
