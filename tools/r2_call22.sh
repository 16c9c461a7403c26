timeout 1500 python -m pytest tests/test_gpu_ops.py -q -x 2>&1 | tail -2
timeout 900 python -m pytest tests/test_gpu_modules.py tests/test_gpu_train.py tests/test_gpu_packplan.py tests/test_gpu_metrics.py -q -x 2>&1 | tail -2
for pdl in 0 1; do
FDG_PDL=$pdl python bench.py --quick --steps 10 --warmup 3 2>&1 | tail -1 | cut -c1-90
FDG_PDL=$pdl B=1 python tools/r2_graph16.py 2>&1 | tail -3
done
