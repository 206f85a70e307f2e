#!/bin/bash
# ncu --set full of one launch of each secondary kernel family (run under gpurun on ONE GPU).
#   bash scripts/profile_kernels.sh r01b "csr compact wide"
# Reports stay in /tmp on the box; only the raw-page CSVs come back (gpurun_out is capped at 64 MiB).
tag=${1:-r01}
what=${2:-"csr compact wide minimizers canon pack8"}
mkdir -p gpurun_out/kernels_${tag}
for w in $what; do
  case $w in
    csr*) re="regex:csr_kernel";; csr_var_compact|csr_compact) re="regex:compact_csr_kernel";; compact*) re="regex:compact_fixed";; pack8|pack64) re="regex:pack_";; minword) re="regex:minimizer_words_kernel";;
    unpack) re="regex:unpack_flat_kernel";; revcomp) re="regex:revcomp_items_kernel";; hist*) re="regex:hist_fixed_kernel";; *) re="regex:fixed_kernel";;
  esac
  timeout 300 ncu --set full --clock-control none --import-source on -k $re -s 2 -c 1 -f -o /tmp/prof_${w} \
      python scripts/prof_one.py $w > gpurun_out/kernels_${tag}/ncu_${w}.log 2>&1
  ncu -i /tmp/prof_${w}.ncu-rep --page raw --csv > gpurun_out/kernels_${tag}/prof_${w}.csv 2>/dev/null
  ncu -i /tmp/prof_${w}.ncu-rep --page source --csv > gpurun_out/kernels_${tag}/src_${w}.csv 2>/dev/null
  tail -1 gpurun_out/kernels_${tag}/ncu_${w}.log
done
ls -la gpurun_out/kernels_${tag}
