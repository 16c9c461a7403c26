set -x
python -m pytest tests/test_gpu_ops.py -x -q -k "test_conv2d" 2>&1 | tail -5 | tee gpurun_out/r2_nt_pytest.log
python tests/ablate.py "D L" 2>&1 | tee gpurun_out/r2_nt_abl.log
FDG_UMMA_NT_FINE=0 python tests/ablate.py "D L" 2>&1 | tee gpurun_out/r2_nt_abl_off.log
