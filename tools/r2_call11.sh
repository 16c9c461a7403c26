set -x
timeout 900 python -m pytest tests/test_gpu_ops.py -q -x -k "split or wgrad" > gpurun_out/r2_c11_ops.log 2>&1
tail -8 gpurun_out/r2_c11_ops.log
timeout 900 python -m pytest tests/test_gpu_modules.py tests/test_gpu_train.py -q -x > gpurun_out/r2_c11_mod.log 2>&1
tail -5 gpurun_out/r2_c11_mod.log
python bench.py --quick --steps 5 --warmup 3 2>&1 | tail -1
