#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; T=${1:-r3l}
timeout 300 python tools/quick_bench.py 0.03 10000 600 1 collagen24 > $O/${T}_cfg4_tmem.log 2>&1
JF_CG_STREAM_TMEM=0 timeout 300 python tools/quick_bench.py 0.03 10000 600 1 collagen24 > $O/${T}_cfg4_notmem.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $O/${T}_tests.log 2>&1; echo "tests rc $?" >> $O/${T}_tests.log
tail -4 $O/${T}_cfg4_tmem.log; tail -4 $O/${T}_cfg4_notmem.log; tail -3 $O/${T}_tests.log
