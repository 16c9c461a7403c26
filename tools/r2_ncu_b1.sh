ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_b1.csv python bench.py --batch 1 --steps 1 --warmup 3 --quick > gpurun_out/r2_ncu_b1.log 2>&1
tail -1 gpurun_out/r2_ncu_b1.log
python bench.py --batch 1 --quick --steps 20 --warmup 5 2>&1 | tail -1
