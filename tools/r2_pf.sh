for pf in 0 4 8 16; do echo "PF=$pf"; FDG_WU_PF=$pf BENCH_GSPLIT=1 python tests/bench_conv.py wgrad "K2 1x1" 2>&1 | sed 's/  */ /g'; done | tee gpurun_out/r2_pf.log
