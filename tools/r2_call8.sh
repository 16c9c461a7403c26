set -x
timeout 900 python -m pytest tests/test_gpu_ops.py -q -x -k "test_wgrad" > gpurun_out/r2_c8_ops.log 2>&1
tail -15 gpurun_out/r2_c8_ops.log
FDG_WGRAD_K1=0 python bench.py --quick --steps 5 --warmup 3 2>&1 | tail -1
FDG_WGRAD_K1=1 python bench.py --quick --steps 5 --warmup 3 2>&1 | tail -1
FDG_WGRAD_K1=0 python tests/bench_conv.py wgrad "K1 3x3" 2>&1 | tail -3
FDG_WGRAD_K1=1 python tests/bench_conv.py wgrad "K1 3x3" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_modules.py tests/test_gpu_train.py -q -x > gpurun_out/r2_c8_mod.log 2>&1
tail -5 gpurun_out/r2_c8_mod.log
