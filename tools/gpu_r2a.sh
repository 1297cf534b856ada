#!/bin/bash
# round-2 first GPU session: baseline tests + bench, CG barrier timing split, CG iteration bias, ncu of element/assembly
cd "$(dirname "$0")/.."
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/r2a_smi.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q -s > $O/r2a_tests.log 2>&1; echo "tests rc $?" >> $O/r2a_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 > $O/r2a_bench.json 2> $O/r2a_bench.err
JF_B200_LIB=$PWD/jointforces_b200/_lib_timing/libjf_b200.so timeout 300 python tools/profile_run.py cfg2 3 > $O/r2a_cgtiming.log 2>&1
timeout 600 python tools/cg_iter_bias.py > $O/r2a_cgbias.log 2>&1
timeout 300 python tools/profile_run.py cfg2 3 > $O/r2a_plain_cfg2.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:jf_element -s 1 -c 1 -o $O/r2a_element_cfg2 python tools/profile_run.py cfg2 3 > $O/r2a_ncu_element.log 2>&1
timeout 600 python tools/profile_run.py cfg5 1 > $O/r2a_plain_cfg5.log 2>&1 && \
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:jf_assemble_K -s 1 -c 1 -o $O/r2a_assemble_cfg5 python tools/profile_run.py cfg5 1 > $O/r2a_ncu_assemble.log 2>&1
ls -la $O | tail -20
