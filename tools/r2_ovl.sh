timeout 300 python -m pytest tests/test_gpu_train.py -x -q 2>&1 | tail -3 | tee gpurun_out/r2_ovl_pytest.log
for f in 1 0 1 0; do FDG_OVERLAP=$f python bench.py --steps 10 --warmup 3 --quick 2>&1 | tail -1 | cut -c1-120; done | tee gpurun_out/r2_ovl.log
