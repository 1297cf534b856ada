#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; T=${1:-r2n2}; N=${2:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 3 --warmup 3 > $O/${T}_bench_n$N.json 2> $O/${T}_bench_n$N.err
echo "rc $?"; head -c 2500 $O/${T}_bench_n$N.json; echo; tail -5 $O/${T}_bench_n$N.err
