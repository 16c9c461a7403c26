timeout 120 python -m pytest tests/test_gpu_ops.py -x -q -k "split_planes or test_wgrad" 2>&1 | tail -5 | tee gpurun_out/r2_xs_pytest.log
ABL_MODES=0 ABL_PLANES=1 timeout 120 python tests/ablate.py "wgrad W" 2>&1 | tee gpurun_out/r2_xs_abl.log
