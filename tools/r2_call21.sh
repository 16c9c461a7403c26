timeout 900 python -m pytest tests/test_gpu_ops.py -q -x -k "test_wgrad" 2>&1 | tail -2
BENCH_B=1 python tests/bench_conv.py wgrad "K1 3x3" 2>&1 | tail -3
python tests/bench_conv.py wgrad "K1 3x3" 2>&1 | tail -3
python bench.py --quick --steps 10 --warmup 3 2>&1 | tail -1 | cut -c1-90
python bench.py --batch 1 --quick --steps 20 --warmup 5 2>&1 | tail -1 | cut -c1-90
timeout 900 python -m pytest tests/test_gpu_modules.py tests/test_gpu_train.py -q -x 2>&1 | tail -2
