#!/bin/bash
for sc in 0.25 1 4; do for w in copy revcomp unpack minword; do python scripts/prof_one.py $w --time --scale $sc | sed "s/^/scale $sc /"; done; done
