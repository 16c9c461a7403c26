set -x
timeout 1200 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_ops.py -q -x -k "test_conv2d or test_wgrad" > gpurun_out/r2_san_memcheck_ops.log 2>&1
tail -4 gpurun_out/r2_san_memcheck_ops.log
timeout 1200 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_packplan.py tests/test_gpu_train.py -q -x -k "pack or train_step_matches" > gpurun_out/r2_san_memcheck_step.log 2>&1
tail -4 gpurun_out/r2_san_memcheck_step.log
timeout 1500 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_ops.py -q -x -k "test_conv2d_1x1_on_channel_slices or test_conv2d_bn_backward_epilogue or test_conv2d_3x3_bn or test_split_bf16 or test_wgrad_wide or test_wgrad_3x3_with_cin or by_taps or (test_wgrad and tcgen05 and (case0 or case12 or case18 or case19 or case20)) or (test_conv2d and tcgen05 and (case0 or case1 or case18 or case25 or case26 or case27 or case28))" > gpurun_out/r2_san_racecheck.log 2>&1
tail -4 gpurun_out/r2_san_racecheck.log
timeout 900 compute-sanitizer --tool synccheck --print-limit 20 python -m pytest tests/test_gpu_ops.py -q -x -k "(test_wgrad and tcgen05 and (case0 or case18)) or (test_conv2d and tcgen05 and (case0 or case1 or case25))" > gpurun_out/r2_san_synccheck.log 2>&1
tail -4 gpurun_out/r2_san_synccheck.log
