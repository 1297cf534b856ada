#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; T=${1:-r3z}
for v in "" p1 p3 "" p1 p3; do
  L=$PWD/jointforces_b200/_lib${v:+_$v}/libjf_b200.so
  echo "== lib ${v:-default(p2)}" >> $O/${T}_prefetch.log
  JF_B200_LIB=$L timeout 300 python tools/pair_probe.py 20 2>&1 | grep "pair, both" >> $O/${T}_prefetch.log
done
cat $O/${T}_prefetch.log
