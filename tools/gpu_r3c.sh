#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; T=${1:-r3c}
for v in "1 0" "1 1" "2 1"; do
  set -- $v
  echo "== lanes $1 tmem $2" >> $O/${T}_probe.log
  JF_TIMING=1 JF_CG_PAIR_LANES=$1 JF_CG_PAIR_TMEM=$2 timeout 120 python tools/pair_probe.py 20 2>&1 | grep "pair\|single\|rror" | sort -u | cut -c1-250 >> $O/${T}_probe.log
  JF_CG_PAIR_LANES=$1 JF_CG_PAIR_TMEM=$2 timeout 300 python -m pytest tests/test_gpu_fullsize.py -x -q -s -k "side_by_side or pair_kernel" 2>&1 | tail -6 | cut -c1-250 >> $O/${T}_probe.log
done
JF_CG_PAIR_LANES=2 JF_CG_PAIR_TMEM=1 JF_B200_LIB=$PWD/jointforces_b200/_lib_timing/libjf_b200.so timeout 120 python tools/profile_pair.py 3 2>&1 | grep "^cgp sys 0" | tail -134 > $O/${T}_timing_l2tmem.log
JF_CG_PAIR_LANES=1 JF_CG_PAIR_TMEM=1 JF_B200_LIB=$PWD/jointforces_b200/_lib_timing/libjf_b200.so timeout 120 python tools/profile_pair.py 3 2>&1 | grep "^cgp sys 0" | tail -134 > $O/${T}_timing_l1tmem.log
cat $O/${T}_probe.log
