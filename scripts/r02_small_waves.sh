#!/bin/bash
# Small waves (few RNG streams per GPU, e.g. cfg-3 sharded over 8 GPUs): which EM execution is fastest?
mkdir -p gpurun_out
run() { # name utrs env...
  name=$1; utrs=$2; shift; shift
  env "$@" timeout 300 python bench.py --utrs $utrs --no-cpu --no-cfg3 --steps 2 --warmup 1 > gpurun_out/sw_$name.json 2> gpurun_out/sw_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/sw_$name.json").read().strip().splitlines()[-1])
    print("$name: value %.0f e2e %.0f waves %d"%(d["value"], d["e2e"]["value"], d["waves_per_step"]), {k: round(v) for k, v in d["phases_alone_ms"].items()})
except Exception as e:
    print("$name FAILED", e); print(open("gpurun_out/sw_$name.err").read()[-500:])
PY
}
for U in 1200 2500; do
  run bsp_$U $U SCAPE_B200_EM=bsp SCAPE_B200_SCAN_TILES=0
  run tail0_$U $U SCAPE_B200_EM=tail SCAPE_B200_TAIL_STEP=0
  run tail8_$U $U SCAPE_B200_EM=tail SCAPE_B200_TAIL_STEP=8
  run tail16_$U $U SCAPE_B200_EM=tail SCAPE_B200_TAIL_STEP=16
  run cluster_$U $U SCAPE_B200_EM=cluster
done
