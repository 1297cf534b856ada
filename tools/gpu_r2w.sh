#!/bin/bash
# round-2 session w: tests, default bench (sweep), reference arm, accumulator-stride A/B, launch list of the bench command
cd "$(dirname "$0")/.."
O=gpurun_out; T=r2w
timeout 1500 python -m pytest tests -m gpu -x -q > $O/${T}_tests.log 2>&1; echo "tests rc $?" >> $O/${T}_tests.log
timeout 900 python bench.py --steps 3 --warmup 3 > $O/${T}_bench.json 2> $O/${T}_bench.err
for s in 1 16 512; do
  JF_CG_ACC_STRIDE=$s timeout 600 python bench.py --workload cfg2 --steps 5 --warmup 3 --no-cpu --no-hbm-leg > $O/${T}_cfg2_stride$s.json 2> $O/${T}_cfg2_stride$s.err
  JF_CG_ACC_STRIDE=$s JF_B200_LIB=$PWD/jointforces_b200/_lib_timing/libjf_b200.so timeout 300 python tools/profile_run.py cfg2 3 > $O/${T}_cgtiming_stride$s.log 2>&1
done
timeout 600 python bench.py --impl reference --steps 1 --cpu-budget 60 > $O/${T}_ref.json 2> $O/${T}_ref.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/${T}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-hbm-leg --no-single > $O/${T}_ncu_bench.log 2>&1
tail -3 $O/${T}_tests.log; head -c 1200 $O/${T}_bench.json; echo; for s in 1 16 512; do head -c 700 $O/${T}_cfg2_stride$s.json | grep -o '"value": [0-9.]*' | head -1; tail -2 $O/${T}_cgtiming_stride$s.log | head -1; done
