timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -3 | tee gpurun_out/r2_perc_pytest.log
for f in 1 0 1 0; do echo "OVERLAP_PERC=$f"; FDG_OVERLAP_PERC=$f python tools/b1_graph.py 2>&1 | tail -1; done | tee gpurun_out/r2_perc.log
for f in 1 0 1 0; do FDG_OVERLAP_PERC=$f python bench.py --steps 10 --warmup 3 --quick 2>&1 | tail -1 | cut -c1-120; done | tee -a gpurun_out/r2_perc.log
