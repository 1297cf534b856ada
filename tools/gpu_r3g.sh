#!/bin/bash
# final evidence of round 2: prefetch A/B on the tensor-memory pair kernel, tests, default bench, reference arm, launch list
cd "$(dirname "$0")/.."
O=gpurun_out; T=${1:-r3g}
for v in "" p3 p4; do
  L=$PWD/jointforces_b200/_lib${v:+_$v}/libjf_b200.so
  echo "== lib ${v:-default(p2)}" >> $O/${T}_prefetch.log
  JF_B200_LIB=$L timeout 300 python tools/pair_probe.py 20 2>&1 | grep "pair, both" >> $O/${T}_prefetch.log
done
timeout 1500 python -m pytest tests -m gpu -x -q > $O/${T}_tests.log 2>&1; echo "tests rc $?" >> $O/${T}_tests.log
timeout 900 python bench.py --steps 3 --warmup 3 > $O/${T}_bench.json 2> $O/${T}_bench.err
timeout 600 python bench.py --impl reference --steps 1 --cpu-budget 60 > $O/${T}_ref.json 2> $O/${T}_ref.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/${T}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-hbm-leg --no-single > $O/${T}_ncu_bench.log 2>&1
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1
cat $O/${T}_prefetch.log; tail -3 $O/${T}_tests.log; head -c 1200 $O/${T}_bench.json; echo; tail -2 $O/${T}_smoke.log
