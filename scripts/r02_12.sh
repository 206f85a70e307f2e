#!/bin/bash
bash scripts/profile_kernels.sh r02h "csr_var compact1" > gpurun_out/r02_prof_h.log 2>&1
tail -2 gpurun_out/r02_prof_h.log
timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q -k "config4" 2>&1 | tail -3
