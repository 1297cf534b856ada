#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; T=${1:-r3i}
for v in j6p2 j4p2 j6p3 j6p1 j2p2; do
  L=$PWD/jointforces_b200/_lib${v:+_$v}/libjf_b200.so
  echo "== lib ${v:-default(j8p2)}" >> $O/${T}_regs.log
  JF_B200_LIB=$L timeout 300 python tools/pair_probe.py 20 2>&1 | grep "pair, both\|rror" >> $O/${T}_regs.log
done
cat $O/${T}_regs.log
