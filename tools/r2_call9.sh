set -x
timeout 900 python -m pytest tests/test_gpu_ops.py -q -x -k "bn_backward" > gpurun_out/r2_c9_ops.log 2>&1
tail -15 gpurun_out/r2_c9_ops.log
timeout 900 python -m pytest tests/test_gpu_modules.py tests/test_gpu_train.py tests/test_gpu_gradcheck.py -q -x > gpurun_out/r2_c9_mod.log 2>&1
tail -5 gpurun_out/r2_c9_mod.log
python bench.py --quick --steps 5 --warmup 3 2>&1 | tail -1
