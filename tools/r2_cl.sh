timeout 120 python -m pytest tests/test_gpu_ops.py -x -q -k "split_planes or test_wgrad" 2>&1 | tail -3 | tee gpurun_out/r2_cl_pytest.log
for cl in 1 0; do echo "CL=$cl"; FDG_WU_CLUSTER=$cl BENCH_GSPLIT=1 timeout 120 python tests/bench_conv.py wgrad "D L4" 2>&1 | sed 's/  */ /g'; done | tee gpurun_out/r2_cl.log
