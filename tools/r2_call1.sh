set -x
nvidia-smi -L
python bench.py --quick --steps 5 --warmup 3 > gpurun_out/r2_c1_base.log 2>&1
FDG_WU_FAST=1 timeout 900 python -m pytest tests -m gpu -q -x -k "wgrad or split or fdgan or train" > gpurun_out/r2_c1_fast_pytest.log 2>&1
tail -3 gpurun_out/r2_c1_fast_pytest.log
FDG_WU_FAST=1 python bench.py --quick --steps 5 --warmup 3 > gpurun_out/r2_c1_fast.log 2>&1
cat gpurun_out/r2_c1_base.log gpurun_out/r2_c1_fast.log | grep quick
timeout 900 compute-sanitizer --tool memcheck --print-limit 30 python -m pytest tests/test_gpu_ops.py -q -x -k "test_conv2d and not simt" > gpurun_out/r2_c1_memcheck_conv.log 2>&1
tail -5 gpurun_out/r2_c1_memcheck_conv.log
timeout 900 compute-sanitizer --tool memcheck --print-limit 30 python -m pytest tests/test_gpu_ops.py -q -x -k "test_wgrad" > gpurun_out/r2_c1_memcheck_wgrad.log 2>&1
tail -5 gpurun_out/r2_c1_memcheck_wgrad.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 30 python -m pytest tests/test_gpu_ops.py -q -x -k "test_conv2d_bn_backward_epilogue or test_split_bf16" > gpurun_out/r2_c1_racecheck.log 2>&1
tail -5 gpurun_out/r2_c1_racecheck.log
