#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; T=${1:-r3t}
bash tools/gpu_final.sh ${T} > $O/${T}_final_script.log 2>&1
timeout 300 python tools/profile_cg_stream.py > $O/${T}_plain_stream.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none -k regex:jf_cg_stream_kernel -s 1 -c 1 -o $O/${T}_cg_stream_large python tools/profile_cg_stream.py > $O/${T}_ncu_stream.log 2>&1
tail -12 $O/${T}_final_script.log | cut -c1-1300; tail -2 $O/${T}_ncu_stream.log
