#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; T=${1:-r3y}
for v in 0 1 2 3 4 5 6; do echo "== JF_ELEM_VARIANT=$v" >> $O/${T}_elem.log; JF_ELEM_VARIANT=$v timeout 120 python tools/quick_bench.py 0.05 100 20 2 2>&1 | grep "rep 1" >> $O/${T}_elem.log; done
for l in 2 8; do echo "== JF_ELEM_LANES=$l" >> $O/${T}_elem.log; JF_ELEM_LANES=$l timeout 120 python tools/quick_bench.py 0.05 100 20 2 2>&1 | grep "rep 1" >> $O/${T}_elem.log; done
cat $O/${T}_elem.log
