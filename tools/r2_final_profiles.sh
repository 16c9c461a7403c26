set -x
ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none --csv --log-file gpurun_out/r02_launches_traffic_final.csv python bench.py --steps 1 --warmup 3 --quick > gpurun_out/r02_ncu_final.log 2>&1
tail -1 gpurun_out/r02_ncu_final.log | cut -c1-200
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_b1_final.csv python bench.py --batch 1 --steps 1 --warmup 3 --quick > gpurun_out/r02_ncu_b1_final.log 2>&1
tail -1 gpurun_out/r02_ncu_b1_final.log | cut -c1-200
ncu --set full --clock-control none --import-source on -k regex:wgrad_k1_kernel -s 1 -c 1 -o gpurun_out/r02_wgrad_k1_128to32 -f python tests/bench_conv.py wgrad "K1 3x3 128->32 @256" > gpurun_out/r02_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_umma_kernel -s 1 -c 1 -o gpurun_out/r02_conv_umma_fast_224to128 -f python tests/bench_conv.py umma "K2 1x1 224" > gpurun_out/r02_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_k1_kernel -s 1 -c 1 -o gpurun_out/r02_conv_k1_128to32 -f python tests/bench_conv.py umma "K1 3x3 128->32 @256" > gpurun_out/r02_ncu3.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:conv_halo_kernel<128, 3, true" -s 1 -c 1 -o gpurun_out/r02_conv_halo_bn2_32to128 -f python tools/bn2_micro.py > gpurun_out/r02_ncu4.log 2>&1
ls -la gpurun_out/*.ncu-rep
bash tools/r2_sanitize.sh > gpurun_out/r02_final_sanitize.log 2>&1; grep -h "SUMMARY\|passed" gpurun_out/r2_san_*.log
