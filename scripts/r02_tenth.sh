#!/bin/bash
bash scripts/profile_kernels.sh r02g "pack64 pack8 revcomp minimizers wide" > gpurun_out/r02_prof_g.log 2>&1
tail -3 gpurun_out/r02_prof_g.log
