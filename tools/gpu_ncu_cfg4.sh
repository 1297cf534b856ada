#!/bin/bash
# ncu --set full of one launch of the streaming single-reduction kernel on config 4 (with and without its on-chip caches)
cd "$(dirname "$0")/.."
T=${1:-ncu4}; O=gpurun_out
timeout 300 python tools/profile_run.py cfg4 2 > $O/${T}_plain.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:jf_cg_fused_stream -s 1 -c 1 -o $O/${T}_cg_fused_stream_cfg4 python tools/profile_run.py cfg4 2 > $O/${T}_ncu.log 2>&1
JF_CG_STREAM_TMEM=0 JF_CG_STREAM_NO_SMEM_CACHE=1 timeout 900 ncu --set full --clock-control none -k regex:jf_cg_fused_stream -s 1 -c 1 -o $O/${T}_cg_fused_stream_cfg4_nocache python tools/profile_run.py cfg4 2 > $O/${T}_ncu_nocache.log 2>&1
tail -2 $O/${T}_plain.log | cut -c1-300; tail -2 $O/${T}_ncu.log; tail -2 $O/${T}_ncu_nocache.log
