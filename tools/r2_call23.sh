timeout 900 python -m pytest tests/test_gpu_modules.py tests/test_gpu_train.py tests/test_gpu_packplan.py tests/test_gpu_gradcheck.py -q -x 2>&1 | tail -2
for px in 0 262144 100000000; do
echo "ASYNC_PIXELS=$px"
FDG_ASYNC_WGRAD_PIXELS=$px python bench.py --quick --steps 10 --warmup 3 2>&1 | tail -1 | cut -c1-90
FDG_ASYNC_WGRAD_PIXELS=$px B=1 python tools/r2_graph16.py 2>&1 | tail -3 | head -2
done
