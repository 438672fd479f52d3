#!/bin/bash
# A/B of environment knobs on the cfg-3 workload.  Usage: scripts/r02_ab3.sh "NAME ENV=.." ...
mkdir -p gpurun_out
for spec in "$@"; do
  set -- $spec; name=$1; shift
  env "$@" timeout 400 python bench.py --workload cfg3 --no-cpu --no-files --steps 1 --warmup 1 > gpurun_out/ab3_$name.json 2> gpurun_out/ab3_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/ab3_$name.json").read().strip().splitlines()[-1])
    print("cfg3 $name: value %.0f e2e %.0f roof %.3f"%(d["value"], d["e2e"]["value"], d["roofline"]["frac"]), {k: round(v) for k, v in d["phases_alone_ms"].items()})
except Exception as e:
    print("$name FAILED", e); print(open("gpurun_out/ab3_$name.err").read()[-600:])
PY
done
