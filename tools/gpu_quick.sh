#!/bin/bash
# quick GPU iteration: parity tests, short bench, per-CTA CG timing split, optional A/B legs.
#   bash tools/gpu_quick.sh TAG [leg ...]    leg = name of a jointforces_b200/_lib_<name> build, or VAR=value (environment switch)
cd "$(dirname "$0")/.."
T=${1:-q}; shift; O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/${T}_tests.log 2>&1; echo "tests rc $?" >> $O/${T}_tests.log
timeout 600 python bench.py --workload cfg2 --steps 5 --warmup 3 --no-cpu --no-hbm-leg > $O/${T}_bench.json 2> $O/${T}_bench.err
JF_B200_LIB=$PWD/jointforces_b200/_lib_timing/libjf_b200.so timeout 300 python tools/profile_run.py cfg2 3 > $O/${T}_cgtiming.log 2>&1
for v in "$@"; do
  if [[ "$v" == *=* ]]; then
    env "$v" timeout 600 python bench.py --workload cfg2 --steps 5 --warmup 3 --no-cpu --no-hbm-leg > $O/${T}_bench_${v%%=*}.json 2> $O/${T}_bench_${v%%=*}.err
    env "$v" JF_B200_LIB=$PWD/jointforces_b200/_lib_timing/libjf_b200.so timeout 300 python tools/profile_run.py cfg2 3 > $O/${T}_cgtiming_${v%%=*}.log 2>&1
  else
    JF_B200_LIB=$PWD/jointforces_b200/_lib_$v/libjf_b200.so timeout 600 python bench.py --workload cfg2 --steps 5 --warmup 3 --no-cpu --no-hbm-leg > $O/${T}_bench_$v.json 2> $O/${T}_bench_$v.err
    [ -d jointforces_b200/_lib_${v}_timing ] && JF_B200_LIB=$PWD/jointforces_b200/_lib_${v}_timing/libjf_b200.so timeout 300 python tools/profile_run.py cfg2 3 > $O/${T}_cgtiming_$v.log 2>&1
  fi
done
tail -3 $O/${T}_tests.log; for f in $O/${T}_bench*.json; do echo $f; head -c 300 $f; echo; done
