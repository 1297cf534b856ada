#!/bin/bash
# final check of a round: GPU tests, default bench, reference arm, smoke
cd "$(dirname "$0")/.."
O=gpurun_out; T=${1:-final}
timeout 1500 python -m pytest tests -m gpu -x -q > $O/${T}_tests.log 2>&1; echo "tests rc $?" >> $O/${T}_tests.log
timeout 900 python bench.py --steps 3 --warmup 3 > $O/${T}_bench.json 2> $O/${T}_bench.err
timeout 600 python bench.py --impl reference --steps 1 --cpu-budget 60 > $O/${T}_ref.json 2> $O/${T}_ref.err
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1
tail -3 $O/${T}_tests.log; head -c 1200 $O/${T}_bench.json; echo; head -c 300 $O/${T}_ref.json; echo; tail -2 $O/${T}_smoke.log | cut -c1-200
