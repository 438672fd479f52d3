#!/bin/bash
# parity subset + A/B lines in one box visit
mkdir -p gpurun_out
: > gpurun_out/exp.txt
source scripts/exp_lib.sh
( timeout 400 python -m pytest tests -m gpu -x -q -k "chain_traces or goldens or ragged or many_streams or heavy_tailed" 2>&1 | tail -5 ) > gpurun_out/pytest_gpu.log
run "A=1" --steps 2 --warmup 2
run "SCAPE_B200_STAGE_CHAIN=0" --steps 2 --warmup 2
C3="--workload cfg3 --utrs 6000 --steps 1 --warmup 1"
run "A=1" $C3
run "SCAPE_B200_STAGE_CHAIN=0" $C3
timeout 200 python scripts/sweep_cfg5.py --budget-s 150 > gpurun_out/sweep_cfg5.jsonl 2> gpurun_out/sweep_cfg5.err
cat gpurun_out/pytest_gpu.log gpurun_out/exp.txt; wc -l gpurun_out/sweep_cfg5.jsonl; tail -3 gpurun_out/sweep_cfg5.err
