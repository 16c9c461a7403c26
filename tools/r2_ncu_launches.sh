set -x
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_a.csv python bench.py --steps 1 --warmup 3 --quick > gpurun_out/r2_ncu_a.log 2>&1
tail -2 gpurun_out/r2_ncu_a.log
python bench.py --quick --steps 5 --warmup 3 2>&1 | tail -1
