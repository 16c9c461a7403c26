timeout 300 python -m pytest tests/test_gpu_ops.py -x -q -k "test_conv2d or bn_backward" 2>&1 | tail -3 | tee gpurun_out/r2_hcl_pytest.log
for cl in 1 0; do echo "CL=$cl"; FDG_HALO_CLUSTER=$cl ABL_MODES=0 timeout 200 python tests/ablate.py "D L,dgrad 3x3,vgg" 2>&1 | grep -v shape; done | tee gpurun_out/r2_hcl.log
