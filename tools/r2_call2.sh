set -x
free -g | head -2; nproc
timeout 2400 python -m pytest tests -m gpu -q -x -s --durations=15 > gpurun_out/r2_c2_pytest.log 2>&1
tail -40 gpurun_out/r2_c2_pytest.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_c2_bench.log 2> gpurun_out/r2_c2_bench.err
cat gpurun_out/r2_c2_bench.log; tail -5 gpurun_out/r2_c2_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_c2_ref.log 2>&1
cat gpurun_out/r2_c2_ref.log
