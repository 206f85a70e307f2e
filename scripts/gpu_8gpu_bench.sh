#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_r02_8gpu_b.json 2> gpurun_out/bench_r02_8gpu_b.err
echo "bench8 rc=$?"; tail -c 600 gpurun_out/bench_r02_8gpu_b.err
