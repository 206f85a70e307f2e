#!/bin/bash
for so in "" exp_so/c3i2304.so exp_so/c2i3072.so exp_so/c2i4096.so; do KMERS_B200_SO=$so python scripts/prof_one.py compact1 --time | sed "s#^#${so:-default2048} #"; done
