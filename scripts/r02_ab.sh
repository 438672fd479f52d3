#!/bin/bash
# A/B of environment knobs on the cfg-2 bench line.  Usage: scripts/r02_ab.sh "NAME ENV=.. ENV=.." ...
mkdir -p gpurun_out
for spec in "$@"; do
  set -- $spec; name=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu --no-cfg3 --no-files --steps 2 --warmup 1 > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/ab_$name.json").read().strip().splitlines()[-1])
    print("$name: value %.0f e2e %.0f roof %.3f"%(d["value"], d["e2e"]["value"], d["roofline"]["frac"]), {k: round(v) for k, v in d["phases_alone_ms"].items()})
except Exception as e:
    print("$name FAILED", e); print(open("gpurun_out/ab_$name.err").read()[-600:])
PY
done
