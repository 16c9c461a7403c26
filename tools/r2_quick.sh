python bench.py --steps 10 --warmup 3 --quick 2>&1 | tail -1 | tee gpurun_out/r2_quick.log
python bench.py --batch 1 --steps 20 --warmup 3 --quick 2>&1 | tail -1 | tee gpurun_out/r2_quick_b1.log
python -m pytest tests/test_gpu_train.py tests/test_gpu_modules.py -x -q 2>&1 | tail -3 | tee gpurun_out/r2_quick_pytest.log
