#!/bin/bash
for so in "" exp_so/packA.so exp_so/packB.so exp_so/packC.so; do
  for w in pack64 pack8; do KMERS_B200_SO=$so python scripts/prof_one.py $w --time --scale 2.5 | sed "s#^#${so:-default} #"; done
done
