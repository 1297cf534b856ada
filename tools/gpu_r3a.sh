#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; T=${1:-r3a}
timeout 120 ./tools/micro/tmem_bench > $O/${T}_tmem_bench.log 2>&1
timeout 300 python tools/pair_probe.py 20 2>&1 | grep "pair\|single" > $O/${T}_probe_smem.log
JF_CG_PAIR_TMEM=1 timeout 300 python tools/pair_probe.py 20 2>&1 | grep "pair\|single" > $O/${T}_probe_tmem.log
JF_CG_PAIR_TMEM=1 timeout 300 python -m pytest tests/test_gpu_fullsize.py -x -q -s -k "side_by_side or pair_kernel" > $O/${T}_pairtests_tmem.log 2>&1
JF_CG_PAIR_TMEM=1 JF_B200_LIB=$PWD/jointforces_b200/_lib_timing/libjf_b200.so timeout 300 python tools/profile_pair.py 3 2>&1 | grep "^cgp sys 0" | tail -134 > $O/${T}_timing_tmem.log
cat $O/${T}_tmem_bench.log; cat $O/${T}_probe_smem.log $O/${T}_probe_tmem.log; tail -4 $O/${T}_pairtests_tmem.log
