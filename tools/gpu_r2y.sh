#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; T=${1:-r2y}
timeout 300 python tools/pair_probe.py 20 > $O/${T}_probe_l1.log 2>&1
JF_CG_PAIR_LANES=2 timeout 300 python tools/pair_probe.py 20 > $O/${T}_probe_l2.log 2>&1
JF_CG_PAIR_LANES=2 timeout 300 python -m pytest tests/test_gpu_fullsize.py -x -q -s -k "side_by_side or pair_kernel" > $O/${T}_pairtests_l2.log 2>&1
JF_B200_LIB=$PWD/jointforces_b200/_lib_timing/libjf_b200.so timeout 300 python tools/pair_probe.py 3 > $O/${T}_timing_l1.log 2>&1
cat $O/${T}_probe_l1.log $O/${T}_probe_l2.log | grep -v Geometry; tail -3 $O/${T}_pairtests_l2.log
