#!/bin/bash
# round-2 session x: first run of jf_cg_pair_kernel (two systems side by side) + cfg4 L2-window A/B
cd "$(dirname "$0")/.."
O=gpurun_out; T=${1:-r2x}
timeout 600 python -m pytest tests/test_gpu_fullsize.py -x -q -s -k "side_by_side or pair_kernel" > $O/${T}_pairtests.log 2>&1; echo "rc $?" >> $O/${T}_pairtests.log
timeout 300 python tools/quick_bench.py 0.05 100 600 1 > $O/${T}_quick_n1.log 2>&1
timeout 300 python tools/quick_bench.py 0.05 100 600 2 > $O/${T}_quick_n2.log 2>&1
JF_B200_LIB=$PWD/jointforces_b200/_lib_timing/libjf_b200.so timeout 300 python tools/quick_bench.py 0.05 100 3 2 > $O/${T}_cgtiming_pair.log 2>&1
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu --no-hbm-leg > $O/${T}_bench.json 2> $O/${T}_bench.err
JF_CG_NO_PAIR=1 timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu --no-hbm-leg --no-single > $O/${T}_bench_nopair.json 2> $O/${T}_bench_nopair.err
timeout 300 python tools/quick_bench.py 0.03 10000 600 1 collagen24 > $O/${T}_cfg4_window.log 2>&1
JF_NO_L2_WINDOW=1 timeout 300 python tools/quick_bench.py 0.03 10000 600 1 collagen24 > $O/${T}_cfg4_nowindow.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $O/${T}_tests.log 2>&1; echo "tests rc $?" >> $O/${T}_tests.log
tail -5 $O/${T}_pairtests.log; tail -4 $O/${T}_quick_n1.log; tail -4 $O/${T}_quick_n2.log; tail -3 $O/${T}_cgtiming_pair.log; head -c 900 $O/${T}_bench.json; echo; tail -3 $O/${T}_tests.log
