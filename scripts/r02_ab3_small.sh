#!/bin/bash
# A/B of environment knobs on a cfg-3 slice (heavy-tailed UTRs, few RNG streams, as one of 8 GPUs sees it).
# Usage: scripts/r02_ab3_small.sh <utrs> "NAME ENV=.." ...
utrs=$1; shift
mkdir -p gpurun_out
for spec in "$@"; do
  set -- $spec; name=$1; shift
  env "$@" timeout 300 python bench.py --workload cfg3 --utrs $utrs --no-cpu --no-files --steps 2 --warmup 1 > gpurun_out/abs3_${utrs}_$name.json 2> gpurun_out/abs3_${utrs}_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/abs3_${utrs}_$name.json").read().strip().splitlines()[-1])
    print("cfg3 utrs $utrs $name: value %.0f e2e %.0f"%(d["value"], d["e2e"]["value"]))
except Exception as e:
    print("$name FAILED", e); print(open("gpurun_out/abs3_${utrs}_$name.err").read()[-600:])
PY
done
