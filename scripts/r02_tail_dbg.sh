#!/bin/bash
mkdir -p gpurun_out
timeout 180 python __graft_entry__.py --smoke 2>&1 | tail -1
for S in ${STEPS:-0 8}; do
  SCAPE_B200_TAIL_STEP=$S timeout 300 python bench.py --no-cpu --no-cfg3 --steps 2 --warmup 1 > gpurun_out/bench_t$S.json 2> gpurun_out/bench_t$S.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_t$S.json").read().strip().splitlines()[-1])
    print("TAIL_STEP=$S cfg2 value %.0f e2e %.0f roof %.3f"%(d["value"], d["e2e"]["value"], d["roofline"]["frac"]), d["phases_alone_ms"])
except Exception as e:
    print("TAIL_STEP=$S FAILED", e); print(open("gpurun_out/bench_t$S.err").read()[-800:])
PY
done
