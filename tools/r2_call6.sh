set -x
timeout 900 python -m pytest tests/test_gpu_ops.py -q -x -k "test_conv2d" > gpurun_out/r2_c6_ops.log 2>&1
tail -5 gpurun_out/r2_c6_ops.log
timeout 900 python -m pytest tests/test_gpu_modules.py tests/test_gpu_train.py tests/test_gpu_packplan.py -q -x > gpurun_out/r2_c6_mod.log 2>&1
tail -5 gpurun_out/r2_c6_mod.log
FDG_CONV_FAST=0 python bench.py --quick --steps 5 --warmup 3 2>&1 | tail -1
FDG_CONV_FAST=1 python bench.py --quick --steps 5 --warmup 3 2>&1 | tail -1
FDG_CONV_FAST=1 FDG_WU_FAST=1 python bench.py --quick --steps 5 --warmup 3 2>&1 | tail -1
FDG_CONV_FAST=0 python tests/bench_conv.py umma 2>&1 | tail -12
FDG_CONV_FAST=1 python tests/bench_conv.py umma 2>&1 | tail -12
