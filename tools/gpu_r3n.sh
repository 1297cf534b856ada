#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; T=${1:-r3n}
timeout 600 python bench.py --workload cfg2 --steps 2 --warmup 3 --no-cpu > $O/${T}_bench_cache.json 2> $O/${T}_bench_cache.err
JF_CG_STREAM_NO_CACHE=1 timeout 600 python bench.py --workload cfg2 --steps 2 --warmup 3 --no-cpu > $O/${T}_bench_nocache.json 2> $O/${T}_bench_nocache.err
timeout 1500 python -m pytest tests -m gpu -x -q -s > $O/${T}_tests.log 2>&1; echo "tests rc $?" >> $O/${T}_tests.log
for f in cache nocache; do python -c "
import json,sys
d=json.loads([l for l in open('$O/${T}_bench_$f.json').read().splitlines() if l.startswith('{')][-1])
print('$f', d.get('roofline_large'))
"; done; grep "^\[cfg5" $O/${T}_tests.log; tail -3 $O/${T}_tests.log
