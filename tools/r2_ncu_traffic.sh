ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none --csv --log-file gpurun_out/r2_launches_traffic.csv python bench.py --steps 1 --warmup 3 --quick > gpurun_out/r2_ncu_t.log 2>&1
tail -1 gpurun_out/r2_ncu_t.log
