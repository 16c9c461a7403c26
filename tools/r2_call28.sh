timeout 900 python -m pytest tests/test_gpu_ops.py -q -x -k "by_taps" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_modules.py tests/test_gpu_train.py -q -x 2>&1 | tail -2
python bench.py --quick --steps 10 --warmup 3 2>&1 | tail -1 | cut -c1-90
