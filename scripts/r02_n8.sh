#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l > gpurun_out/r02_n8_gpus.txt; nproc >> gpurun_out/r02_n8_gpus.txt; free -g | head -2 >> gpurun_out/r02_n8_gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_r02_8gpu.json 2> gpurun_out/bench_r02_8gpu.err
echo "bench8 rc=$?"; tail -c 800 gpurun_out/bench_r02_8gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/bench_r02_4gpu.json 2> gpurun_out/bench_r02_4gpu.err
echo "bench4 rc=$?"
