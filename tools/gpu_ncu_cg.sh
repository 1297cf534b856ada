#!/bin/bash
# ncu --set full of one CG launch of the current kernel on config 2 (3 Newton steps; the second CG launch is captured)
cd "$(dirname "$0")/.."
T=${1:-ncu}; O=gpurun_out
timeout 300 python tools/profile_run.py cfg2 3 > $O/${T}_plain.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:jf_cg_fused -s 1 -c 1 -o $O/${T}_cg_fused_cfg2 python tools/profile_run.py cfg2 3 > $O/${T}_ncu.log 2>&1
tail -2 $O/${T}_plain.log; tail -3 $O/${T}_ncu.log
