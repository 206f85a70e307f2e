#!/bin/bash
# 2-GPU validation: NCCL tests, bench under torchrun, compaction re-test
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_n2_gpus.txt
timeout 600 python -m pytest tests/test_gpu_dist.py tests/test_gpu_compact.py -m gpu -q -rs > gpurun_out/pytest_2gpu_r02.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_2gpu_r02.log
tail -4 gpurun_out/pytest_2gpu_r02.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_r02_2gpu.json 2> gpurun_out/bench_r02_2gpu.err
echo "bench2 rc=$?"; tail -c 1500 gpurun_out/bench_r02_2gpu.err
python scripts/prof_one.py compact1 --time
