set -x
timeout 900 python -m pytest tests/test_gpu_ops.py -q -x -k "test_conv2d" > gpurun_out/r2_c17_ops.log 2>&1
tail -4 gpurun_out/r2_c17_ops.log
python tests/bench_conv.py umma "K2 1x1" 2>&1 | tail -4
python tests/bench_conv.py umma "dgrad 1x1" 2>&1 | tail -1
python bench.py --quick --steps 10 --warmup 3 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_modules.py tests/test_gpu_train.py -q -x > gpurun_out/r2_c17_mod.log 2>&1
tail -3 gpurun_out/r2_c17_mod.log
