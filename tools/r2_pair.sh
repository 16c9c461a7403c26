python -m pytest tests/test_gpu_ops.py -x -q -k "test_conv2d" 2>&1 | tail -3 | tee gpurun_out/r2_pair_pytest.log
ABL_MODES=0,4,15 python tests/ablate.py "dgrad 3x3" 2>&1 | tee gpurun_out/r2_pair_abl.log
FDG_UMMA_TAP_PAIR=0 ABL_MODES=0,4,15 python tests/ablate.py "dgrad 3x3" 2>&1 | tee gpurun_out/r2_pair_abl_off.log
