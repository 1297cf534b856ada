#!/bin/bash
# ncu --set full of one launch of jf_cg_pair_kernel (two systems side by side, config 2; the second CG launch is captured)
cd "$(dirname "$0")/.."
T=${1:-ncu}; O=gpurun_out
timeout 300 python tools/profile_pair.py 3 > $O/${T}_plain.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:jf_cg_pair -s 1 -c 1 -o $O/${T}_cg_pair_cfg2 python tools/profile_pair.py 3 > $O/${T}_ncu.log 2>&1
tail -2 $O/${T}_plain.log | cut -c1-400; tail -3 $O/${T}_ncu.log
