#!/bin/bash
# A/B template for one box visit: parity subset, then bench lines with environment knobs;
# one summary line per run in gpurun_out/exp.txt (see scripts/exp_lib.sh).
# NB when wrapping gpurun in a retry loop: test for "status=transient", never for the word "busy"
# (a successful bench line contains "device_busy_ms").
mkdir -p gpurun_out
: > gpurun_out/exp.txt
source scripts/exp_lib.sh
( timeout 400 python -m pytest tests -m gpu -x -q -k "chain_traces or goldens or ragged or many_streams or heavy_tailed" 2>&1 | tail -5 ) > gpurun_out/pytest_gpu.log
run "A=1" --steps 2 --warmup 2
run "SCAPE_B200_PREDRAW=0" --steps 2 --warmup 2
C3="--workload cfg3 --utrs 6000 --steps 1 --warmup 1"
run "A=1" $C3
cat gpurun_out/pytest_gpu.log gpurun_out/exp.txt
