set -x
timeout 1500 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_ops.py -q -x -k "test_conv2d_1x1_on_channel_slices or test_conv2d_bn_backward_epilogue or test_conv2d_3x3_bn or test_split_bf16 or test_wgrad_wide or test_wgrad_3x3_with_cin or by_taps or (test_wgrad and tcgen05 and (case0 or case12 or case18 or case19 or case20)) or (test_conv2d and tcgen05 and (case0 or case1 or case18 or case25 or case26 or case27 or case28))" > gpurun_out/r2_san_racecheck.log 2>&1
tail -3 gpurun_out/r2_san_racecheck.log
timeout 600 python -m pytest tests/test_gpu_ops.py -q -x -k "wgrad" 2>&1 | tail -1
bash tools/r2_final_profiles.sh > gpurun_out/r02_final_profiles.log 2>&1
tail -3 gpurun_out/r02_final_profiles.log
