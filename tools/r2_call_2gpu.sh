set -x
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_dataparallel.py -q -s > gpurun_out/r2_2gpu_dp_pytest.log 2>&1
tail -5 gpurun_out/r2_2gpu_dp_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_2gpu_bench.log 2> gpurun_out/r2_2gpu_bench.err
tail -1 gpurun_out/r2_2gpu_bench.log; tail -5 gpurun_out/r2_2gpu_bench.err
