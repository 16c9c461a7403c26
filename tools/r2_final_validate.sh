set -x
timeout 3000 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/r02f_pytest.log 2>&1
tail -12 gpurun_out/r02f_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02f_ref.log 2>&1; cut -c1-400 gpurun_out/r02f_ref.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r02f_bench.log 2> gpurun_out/r02f_bench.err
cut -c1-300 gpurun_out/r02f_bench.log; tail -3 gpurun_out/r02f_bench.err
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:conv_halo_kernel<.int.128, .int.3, .bool.1" -s 1 -c 1 -o gpurun_out/r02_conv_halo_bn2_32to128 -f python tools/bn2_micro.py > gpurun_out/r02_ncu4.log 2>&1
ls -la gpurun_out/r02_conv_halo_bn2_32to128.ncu-rep
