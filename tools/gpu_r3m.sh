#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; T=${1:-r3m}
timeout 300 python tools/quick_bench.py 0.03 10000 600 1 collagen24 > $O/${T}_cfg4_tmem_smem.log 2>&1
JF_CG_STREAM_NO_SMEM_CACHE=1 timeout 300 python tools/quick_bench.py 0.03 10000 600 1 collagen24 > $O/${T}_cfg4_tmem.log 2>&1
JF_CG_STREAM_TMEM=0 timeout 300 python tools/quick_bench.py 0.03 10000 600 1 collagen24 > $O/${T}_cfg4_smem.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $O/${T}_tests.log 2>&1; echo "tests rc $?" >> $O/${T}_tests.log
timeout 600 python bench.py --workload cfg4 --steps 3 --warmup 3 --no-cpu --no-hbm-leg > $O/${T}_bench_cfg4.json 2> $O/${T}_bench_cfg4.err
for f in tmem_smem tmem smem; do echo $f; tail -3 $O/${T}_cfg4_$f.log | head -2; done; tail -3 $O/${T}_tests.log; head -c 600 $O/${T}_bench_cfg4.json
