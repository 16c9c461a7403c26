timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -k "bn_backward" 2>&1 | tail -2 | tee gpurun_out/r2_hcl2_pytest.log
timeout 900 python -m pytest tests/test_gpu_modules.py tests/test_gpu_train.py tests/test_gpu_gradcheck.py -x -q 2>&1 | tail -2 | tee -a gpurun_out/r2_hcl2_pytest.log
for f in 1 0 1 0; do FDG_HALO_CLUSTER=$f timeout 300 python bench.py --steps 10 --warmup 3 --quick 2>&1 | tail -1 | cut -c1-120; done | tee gpurun_out/r2_hcl2.log
for f in 1 0; do echo "HALO_CLUSTER=$f"; FDG_HALO_CLUSTER=$f timeout 300 python tools/b1_graph.py 2>&1 | tail -1; done | tee -a gpurun_out/r2_hcl2.log
