timeout 900 python -m pytest tests/test_gpu_ops.py -q -x -k "pool2" 2>&1 | tail -2
timeout 1200 python -m pytest tests/test_gpu_modules.py tests/test_gpu_train.py tests/test_gpu_fullsize.py -q -x 2>&1 | tail -2
python bench.py --quick --steps 10 --warmup 3 2>&1 | tail -1 | cut -c1-90
