#!/bin/bash
# final evidence of round 2 on ONE GPU: launch list + ncu --set full of the bench kernel, per-config timings, GPU test log
mkdir -p gpurun_out
bash scripts/profile.sh r02 > gpurun_out/r02_profile_sh.log 2>&1
tail -3 gpurun_out/r02_profile_sh.log
timeout 900 python scripts/bench_configs.py --out gpurun_out/configs_r02.json > gpurun_out/configs_r02.log 2>&1
tail -45 gpurun_out/configs_r02.log
