set -x
for pf in 0 2 4 8; do echo "PF=$pf"; FDG_CONV_PF=$pf python tests/bench_conv.py umma "K2 1x1" 2>&1 | tail -4; FDG_CONV_PF=$pf python tests/bench_conv.py umma "dgrad 1x1" 2>&1 | tail -1; done
timeout 900 python -m pytest tests/test_gpu_ops.py -q -x -k "test_conv2d" 2>&1 | tail -2
FDG_CONV_PF=0 python bench.py --quick --steps 10 --warmup 3 2>&1 | tail -1
FDG_CONV_PF=4 python bench.py --quick --steps 10 --warmup 3 2>&1 | tail -1
