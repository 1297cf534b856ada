#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; T=${1:-r3s}
for n in 1 3 4; do echo "== JF_CG_STREAM_CACHE=1 JF_CG_STREAM_CACHE_SMEM=$n" >> $O/${T}_hbm.log; JF_CG_STREAM_CACHE=1 JF_CG_STREAM_CACHE_SMEM=$n timeout 300 python tools/hbm_leg.py 2>/dev/null | tail -1 >> $O/${T}_hbm.log; done
grep -o '== .*\|"us_per_cg_iteration": [0-9.]*' $O/${T}_hbm.log
