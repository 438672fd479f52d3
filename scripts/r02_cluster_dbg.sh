#!/bin/bash
# In-kernel phase timing of the cluster-resident EM kernel (development aid) + bench lines for a few knob settings.
mkdir -p gpurun_out
timeout 180 python __graft_entry__.py --smoke 2>&1 | tail -1
for S in 0 4 8 16; do
  SCAPE_B200_DBG=1 SCAPE_B200_OVERLAP=0 SCAPE_B200_SOLO=$S timeout 200 python bench.py --utrs 3000 --steps 1 --warmup 0 --no-cpu --no-cfg3 > gpurun_out/dbg5_s$S.json 2> gpurun_out/dbg5_s$S.err
  echo "== SOLO=$S"; grep "cluster EM" gpurun_out/dbg5_s$S.err | tail -2 | cut -c1-400
done
for S in "4 8" "8 8" "16 8" "8 4"; do
  set -- $S
  SCAPE_B200_SOLO=$1 SCAPE_B200_CLUSTER=$2 timeout 300 python bench.py --no-cpu --no-cfg3 --steps 2 --warmup 1 > gpurun_out/bench_s$1c$2.json 2> gpurun_out/bench_s$1c$2.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_s$1c$2.json").read().strip().splitlines()[-1])
print("SOLO=$1 C=$2 cfg2 value %.0f e2e %.0f roof %.3f"%(d["value"], d["e2e"]["value"], d["roofline"]["frac"]), d["phases_alone_ms"])
PY
done
