#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r02_final.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu_r02_final.log
for w in minimizers unpack minword pack64 pack8; do python scripts/prof_one.py $w --time | cut -c1-150; done
