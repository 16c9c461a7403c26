python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r2_full_pytest.log
python bench.py --steps 10 --warmup 3 --quick 2>&1 | tail -1 | tee gpurun_out/r2_quick.log
python bench.py --batch 1 --steps 20 --warmup 3 --quick 2>&1 | tail -1 | tee gpurun_out/r2_quick_b1.log
