set -x
# one --set full capture per round-2 kernel, on the tests/bench_conv.py shapes (batch 16)
ncu --set full --clock-control none --import-source on -k regex:conv_umma_kernel -s 1 -c 1 -o gpurun_out/r02_conv_umma_fast_224to128 -f python tests/bench_conv.py umma "K2 1x1 224" > gpurun_out/r02_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:wgrad_k1_kernel -s 1 -c 1 -o gpurun_out/r02_wgrad_k1_128to32 -f python tests/bench_conv.py wgrad "K1 3x3 128->32 @256" > gpurun_out/r02_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_k1_kernel -s 1 -c 1 -o gpurun_out/r02_conv_k1_128to32 -f python tests/bench_conv.py umma "K1 3x3 128->32 @256" > gpurun_out/r02_ncu3.log 2>&1
ls -la gpurun_out/*.ncu-rep
