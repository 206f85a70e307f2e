#!/bin/bash
for w in pack64 pack8; do python scripts/prof_one.py $w --time; done
bash scripts/profile_kernels.sh r02l "pack64"
