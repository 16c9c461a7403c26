python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/r2_full_pytest.log
python bench.py --steps 10 --warmup 3 --quick 2>&1 | tail -1 | tee gpurun_out/r2_quick.log
