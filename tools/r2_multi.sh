set -x
N=$1
nvidia-smi -L | wc -l
if [ "$N" = "2" ]; then timeout 600 python -m pytest tests/test_gpu_dataparallel.py -q -s > gpurun_out/r02_${N}gpu_dp_pytest.log 2>&1; tail -3 gpurun_out/r02_${N}gpu_dp_pytest.log; fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r02_${N}gpu_bench.log 2> gpurun_out/r02_${N}gpu_bench.err
tail -1 gpurun_out/r02_${N}gpu_bench.log | cut -c1-300; tail -3 gpurun_out/r02_${N}gpu_bench.err
