#!/bin/bash
# 2-GPU validation of the final build: NCCL tests, both bench arms under torchrun exactly as the driver launches them
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -q -rs > gpurun_out/pytest_2gpu_r02.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_2gpu_r02.log; tail -3 gpurun_out/pytest_2gpu_r02.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/bench_r02_2gpu_ref.json 2> gpurun_out/bench_r02_2gpu_ref.err; echo "ref rc=$?"; head -c 250 gpurun_out/bench_r02_2gpu_ref.json; echo
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_r02_2gpu.json 2> gpurun_out/bench_r02_2gpu.err; echo "bench2 rc=$?"; tail -c 300 gpurun_out/bench_r02_2gpu.err; head -c 250 gpurun_out/bench_r02_2gpu.json
