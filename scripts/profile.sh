#!/bin/bash
# ncu evidence for the bench kernel (run under gpurun on ONE GPU).  usage: bash scripts/profile.sh r01
# Writes gpurun_out/launches_<tag>.csv and gpurun_out/prof_<tag>.ncu-rep ; summarise here with
#   python scripts/summarize_ncu.py <tag>
tag=${1:-r01}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-configs > gpurun_out/ncu_launches_${tag}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fixed_kernel -s 4 -c 1 -f -o /tmp/prof_${tag} \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-configs > gpurun_out/ncu_full_${tag}.log 2>&1
tail -2 gpurun_out/ncu_full_${tag}.log; ncu -i /tmp/prof_${tag}.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_raw.csv 2>/dev/null; ncu -i /tmp/prof_${tag}.ncu-rep --page details > gpurun_out/prof_${tag}_details.txt 2>/dev/null
