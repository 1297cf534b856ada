#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; T=${1:-r3b}
for v in "1 1" "2 1"; do
  set -- $v
  echo "== lanes $1 tmem $2" >> $O/${T}_diag.log
  JF_TIMING=1 JF_CG_PAIR_LANES=$1 JF_CG_PAIR_TMEM=$2 timeout 300 python tools/profile_pair.py 3 2>&1 | grep -i "pair_configure\|^pair\|^newton\|rror" | cut -c1-300 >> $O/${T}_diag.log
done
cat $O/${T}_diag.log
