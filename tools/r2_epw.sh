timeout 400 python -m pytest tests/test_gpu_ops.py -x -q -k "test_conv2d or bn_backward" 2>&1 | tail -3 | tee gpurun_out/r2_epw_pytest.log
for w in 1 0; do echo "EPI_WARP=$w"; FDG_EPI_WARP=$w ABL_MODES=0 timeout 200 python tests/ablate.py "K2 1x1,dgrad 3x3,vgg 3x3 64,D L3,refine4" 2>&1 | grep -v shape; done | tee gpurun_out/r2_epw.log
for w in 1 0 1 0; do FDG_EPI_WARP=$w python bench.py --steps 10 --warmup 3 --quick 2>&1 | tail -1 | cut -c1-120; done | tee -a gpurun_out/r2_epw.log
FDG_ASYNC_WGRAD_PIXELS=100000000 python bench.py --steps 10 --warmup 3 --quick 2>&1 | tail -1 | cut -c1-120 | tee -a gpurun_out/r2_epw.log
