timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -k "bn_backward or test_conv2d" 2>&1 | tail -2 | tee gpurun_out/r2_bn2c_pytest.log
FDG_FUSED_BN2_BWD=1 timeout 900 python -m pytest tests/test_gpu_modules.py tests/test_gpu_train.py tests/test_gpu_gradcheck.py -x -q 2>&1 | tail -2 | tee -a gpurun_out/r2_bn2c_pytest.log
for f in 1 0 1 0; do FDG_FUSED_BN2_BWD=$f python bench.py --steps 10 --warmup 3 --quick 2>&1 | tail -1 | cut -c1-120; done | tee gpurun_out/r2_bn2c.log
for f in 1 0; do echo "FUSED_BN2=$f"; FDG_FUSED_BN2_BWD=$f python tools/b1_graph.py 2>&1 | tail -1; done | tee -a gpurun_out/r2_bn2c.log
