timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -k "test_conv2d or bn_backward" 2>&1 | tail -2 | tee gpurun_out/r2_hcl3_pytest.log
for cl in 1 0; do echo "CL=$cl"; FDG_HALO_CLUSTER=$cl ABL_MODES=0 timeout 200 python tests/ablate.py "D L,dgrad 3x3,vgg" 2>&1 | grep -v "shape\|wgrad"; done | tee gpurun_out/r2_hcl3.log
for f in 1 0 1 0; do FDG_HALO_CLUSTER=$f timeout 300 python bench.py --steps 10 --warmup 3 --quick 2>&1 | tail -1 | cut -c1-120; done | tee -a gpurun_out/r2_hcl3.log
