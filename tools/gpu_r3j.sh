#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; T=${1:-r3j}
timeout 1500 python -m pytest tests -m gpu -x -q > $O/${T}_tests.log 2>&1; echo "tests rc $?" >> $O/${T}_tests.log
timeout 900 python bench.py --steps 3 --warmup 3 > $O/${T}_bench.json 2> $O/${T}_bench.err
timeout 300 python tools/pair_probe.py 20 2>&1 | grep "pair\|single" > $O/${T}_probe.log
bash tools/gpu_ncu_pair.sh ${T} > $O/${T}_ncu_script.log 2>&1
tail -3 $O/${T}_tests.log; head -c 1000 $O/${T}_bench.json; echo; cat $O/${T}_probe.log
