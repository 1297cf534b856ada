#!/bin/bash
# round-2 session 3d: tensor-memory tail as the default of the pair kernel: full tests, default bench, ncu capture
cd "$(dirname "$0")/.."
O=gpurun_out; T=${1:-r3d}
timeout 1500 python -m pytest tests -m gpu -x -q > $O/${T}_tests.log 2>&1; echo "tests rc $?" >> $O/${T}_tests.log
timeout 900 python bench.py --steps 3 --warmup 3 > $O/${T}_bench.json 2> $O/${T}_bench.err
bash tools/gpu_ncu_pair.sh ${T} > $O/${T}_ncu_script.log 2>&1
tail -3 $O/${T}_tests.log; head -c 1500 $O/${T}_bench.json; echo; tail -3 $O/${T}_ncu_script.log
