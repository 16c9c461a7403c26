set -x
timeout 3000 python -m pytest tests -m gpu -q -x --durations=8 > gpurun_out/r2_full_pytest.log 2>&1
tail -14 gpurun_out/r2_full_pytest.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r2_full_bench.log 2> gpurun_out/r2_full_bench.err
cat gpurun_out/r2_full_bench.log; tail -3 gpurun_out/r2_full_bench.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
