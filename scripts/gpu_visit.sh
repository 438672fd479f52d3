#!/bin/bash
# parity subset + A/B lines in one box visit
mkdir -p gpurun_out
: > gpurun_out/exp.txt
source scripts/exp_lib.sh
( timeout 400 python -m pytest tests -m gpu -x -q -k "chain_traces or goldens or ragged or many_streams or heavy_tailed" 2>&1 | tail -5 ) > gpurun_out/pytest_gpu.log
run "A=1" --steps 2 --warmup 2
run "SCAPE_B200_WARP_PF=1" --steps 2 --warmup 2
C3="--workload cfg3 --utrs 6000 --steps 1 --warmup 1"
run "A=1" $C3
run "SCAPE_B200_WARP_PF=1" $C3
SCAPE_B200_WARP_PF=1 SCAPE_B200_OVERLAP=0 SCAPE_B200_DBG=1 timeout 100 python bench.py --utrs 100 --per-file 1 --steps 1 --warmup 1 --no-cpu 2>&1 | grep -A2 "em run" | tail -6 > gpurun_out/dbg_cfg2.txt
cat gpurun_out/pytest_gpu.log gpurun_out/exp.txt; cut -c1-600 gpurun_out/dbg_cfg2.txt
