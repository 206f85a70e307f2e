#!/bin/bash
for cfg in "1536 2304" "1792 2304" "2016 2304" "2016 2816" "2560 2304" "2560 3328" "1280 1792"; do set -- $cfg; for w in csr_var csr; do KMB_CSR_IPC=$1 KMB_CSR_TILE=$2 python scripts/prof_one.py $w --time | sed "s#^#ipc $1 tile $2 #" | cut -c1-120; done; done
