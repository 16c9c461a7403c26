timeout 1200 python -m pytest tests/test_gpu_modules.py tests/test_gpu_train.py tests/test_gpu_gradcheck.py tests/test_gpu_packplan.py -q -x 2>&1 | tail -2
python bench.py --quick --steps 10 --warmup 3 2>&1 | tail -1 | cut -c1-90
B=1 python tools/r2_graph16.py 2>&1 | tail -3 | head -2
B=4 python tools/r2_graph16.py 2>&1 | tail -3 | head -2
