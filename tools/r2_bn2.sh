for f in 0 1 0 1; do FDG_FUSED_BN2_BWD=$f python bench.py --steps 10 --warmup 3 --quick 2>&1 | tail -1 | cut -c1-230; done | tee gpurun_out/r2_bn2.log
