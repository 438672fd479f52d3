#!/bin/bash
# A/B of environment knobs on small cfg-2-shaped workloads (few RNG streams per GPU, as on 8 GPUs).
# Usage: scripts/r02_ab_small.sh <utrs> "NAME ENV=.." ...
utrs=$1; shift
mkdir -p gpurun_out
for spec in "$@"; do
  set -- $spec; name=$1; shift
  env "$@" timeout 300 python bench.py --utrs $utrs --no-cpu --no-cfg3 --no-files --steps 3 --warmup 2 > gpurun_out/abs_${utrs}_$name.json 2> gpurun_out/abs_${utrs}_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/abs_${utrs}_$name.json").read().strip().splitlines()[-1])
    print("utrs $utrs $name: value %.0f e2e %.0f"%(d["value"], d["e2e"]["value"]))
except Exception as e:
    print("$name FAILED", e); print(open("gpurun_out/abs_${utrs}_$name.err").read()[-600:])
PY
done
