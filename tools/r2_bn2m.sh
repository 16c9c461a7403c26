timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -k "bn_backward" 2>&1 | tail -2
for f in 1 0; do FDG_HALO_CLUSTER=$f timeout 200 python tools/bn2_micro.py 2>&1 | tail -3; done | tee gpurun_out/r2_bn2m.log
for f in 1 0 1 0; do FDG_FUSED_BN2_BWD=$f timeout 300 python bench.py --steps 10 --warmup 3 --quick 2>&1 | tail -1 | cut -c1-120; done | tee -a gpurun_out/r2_bn2m.log
