mkdir -p gpurun_out/kernels_r01m
timeout 300 ncu --set full --clock-control none --import-source on -k regex:compact_fixed_kernel -s 4 -c 2 -f -o /tmp/prof_cc python scripts/prof_one.py compact > gpurun_out/kernels_r01m/ncu_cc.log 2>&1
ncu -i /tmp/prof_cc.ncu-rep --page raw --csv > gpurun_out/kernels_r01m/prof_compact2.csv 2>/dev/null
ncu -i /tmp/prof_cc.ncu-rep --page source --csv > gpurun_out/kernels_r01m/src_compact2.csv 2>/dev/null
tail -1 gpurun_out/kernels_r01m/ncu_cc.log
