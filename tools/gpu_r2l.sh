#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2l_tests.log 2>&1; echo "tests rc $?" >> $O/r2l_tests.log
timeout 800 python bench.py --steps 3 --warmup 3 > $O/r2l_bench_n1.json 2> $O/r2l_bench_n1.err
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 > $O/r2l_bench_n2.json 2> $O/r2l_bench_n2.err
timeout 600 python bench.py --impl reference --steps 2 --cpu-budget 60 > $O/r2l_ref.json 2> $O/r2l_ref.err
tail -3 $O/r2l_tests.log; head -c 1500 $O/r2l_bench_n1.json; echo; head -c 600 $O/r2l_bench_n2.json; echo; tail -3 $O/r2l_bench_n2.err
