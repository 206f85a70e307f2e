#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/bench_r02_4gpu.json 2> gpurun_out/bench_r02_4gpu.err
echo "bench4 rc=$?"; tail -c 300 gpurun_out/bench_r02_4gpu.err
