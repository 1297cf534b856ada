#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; T=${1:-r3a}
timeout 120 ./tools/micro/tmem_bench > $O/${T}_tmem_bench.log 2>&1
for v in "1 0" "1 1" "2 1"; do
  set -- $v
  echo "== lanes $1 tmem $2" >> $O/${T}_probe.log
  JF_CG_PAIR_LANES=$1 JF_CG_PAIR_TMEM=$2 timeout 300 python tools/pair_probe.py 20 2>&1 | grep "pair\|single\|rror" >> $O/${T}_probe.log
  JF_CG_PAIR_LANES=$1 JF_CG_PAIR_TMEM=$2 timeout 300 python -m pytest tests/test_gpu_fullsize.py -x -q -s -k "side_by_side or pair_kernel" 2>&1 | tail -4 >> $O/${T}_probe.log
done
JF_CG_PAIR_LANES=2 JF_CG_PAIR_TMEM=1 JF_B200_LIB=$PWD/jointforces_b200/_lib_timing/libjf_b200.so timeout 300 python tools/profile_pair.py 3 2>&1 | grep "^cgp sys 0" | tail -134 > $O/${T}_timing_l2tmem.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu --no-hbm-leg --no-single > $O/${T}_bench.json 2> $O/${T}_bench.err
cat $O/${T}_tmem_bench.log; cat $O/${T}_probe.log; head -c 600 $O/${T}_bench.json
