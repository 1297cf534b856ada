#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; T=${1:-r3r}
echo "== default" >> $O/${T}_hbm.log; timeout 300 python tools/hbm_leg.py 2>/dev/null | tail -1 >> $O/${T}_hbm.log
echo "== JF_CG_STREAM_CACHE=1 (tensor memory only)" >> $O/${T}_hbm.log; JF_CG_STREAM_CACHE=1 timeout 300 python tools/hbm_leg.py 2>/dev/null | tail -1 >> $O/${T}_hbm.log
echo "== JF_CG_STREAM_CACHE=1 JF_CG_STREAM_CACHE_SMEM=5" >> $O/${T}_hbm.log; JF_CG_STREAM_CACHE=1 JF_CG_STREAM_CACHE_SMEM=5 timeout 300 python tools/hbm_leg.py 2>/dev/null | tail -1 >> $O/${T}_hbm.log
echo "== JF_CG_STREAM_CACHE=1 JF_CG_STREAM_CACHE_SMEM=2" >> $O/${T}_hbm.log; JF_CG_STREAM_CACHE=1 JF_CG_STREAM_CACHE_SMEM=2 timeout 300 python tools/hbm_leg.py 2>/dev/null | tail -1 >> $O/${T}_hbm.log
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "on_chip_caches" > $O/${T}_tests.log 2>&1; tail -2 $O/${T}_tests.log
cut -c1-330 $O/${T}_hbm.log
