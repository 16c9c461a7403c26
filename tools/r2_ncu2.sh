ncu --set full --clock-control none --import-source on -k regex:wgrad_umma_kernel -s 1 -c 1 -o gpurun_out/r02_wgrad_umma_dl4 -f python tests/bench_conv.py wgrad "D L4" > gpurun_out/r02_ncu4.log 2>&1
tail -2 gpurun_out/r02_ncu4.log
ncu --set full --clock-control none --import-source on -k regex:conv_halo_kernel -s 1 -c 1 -o gpurun_out/r02_conv_halo_dgrad32 -f python tests/bench_conv.py umma "dgrad 3x3" > gpurun_out/r02_ncu5.log 2>&1
tail -2 gpurun_out/r02_ncu5.log
