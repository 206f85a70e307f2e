#!/bin/bash
for so in exp_so/v16.so exp_so/v12.so exp_so/v32.so; do for w in unpack minword pack64 pack8; do KMERS_B200_SO=$so python scripts/prof_one.py $w --time --scale 2 | sed "s#^#${so:-default} #" | cut -c1-160; done; done
