#!/bin/bash
bash scripts/profile_kernels.sh r02n "compact1" 2>&1 | tail -1
