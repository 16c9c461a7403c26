timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -k "test_conv2d or bn_backward" 2>&1 | tail -2 | tee gpurun_out/r2_hpf_pytest.log
for pf in 1 0; do echo "PF=$pf"; FDG_HALO_PF=$pf ABL_MODES=0 timeout 200 python tests/ablate.py "D L,dgrad 3x3,vgg" 2>&1 | grep -v "shape\|wgrad"; FDG_HALO_PF=$pf timeout 200 python tools/bn2_micro.py 2>&1 | tail -3; done | tee gpurun_out/r2_hpf.log
for f in 1 0 1 0; do FDG_HALO_PF=$f timeout 300 python bench.py --steps 10 --warmup 3 --quick 2>&1 | tail -1 | cut -c1-120; done | tee -a gpurun_out/r2_hpf.log
