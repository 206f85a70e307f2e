#!/bin/bash
for ipc in 1024 1536 2048; do for w in full minimizers wide csr_min csr_wide canon; do KMB_IPC_X=$ipc python scripts/prof_one.py $w --time | sed "s#^#ipc $ipc #" | cut -c1-150; done; done
