#!/bin/bash
# development aid: sweep the EM step at which the next wave's likelihood phase is released / lanes
run() {
  env "$@" timeout 200 python bench.py --steps 2 --warmup 2 --no-cpu > gpurun_out/sw.log 2>&1
  python - "$*" <<PY
import json,sys
d=json.loads(open("gpurun_out/sw.log").read().strip().splitlines()[-1])
p=d["phases_ms_per_step"]
print(sys.argv[1], "value %.0f e2e %.0f busy %.0f total %.0f em %.0f tensor %.0f rng %.0f" % (d["value"], d["e2e"]["value"], p["device_busy_ms"], p["total_ms"], p["em_ms"], p["tensor_ms"], p["host_rng_ms"]))
PY
}
run SCAPE_B200_STAGE_STEP=32
run SCAPE_B200_STAGE_STEP=40
run SCAPE_B200_STAGE_STEP=48
run SCAPE_B200_STAGE_STEP=32 SCAPE_B200_LANES=2
run SCAPE_B200_STAGE_STEP=-1 SCAPE_B200_LANES=2
run SCAPE_B200_OVERLAP=0 SCAPE_B200_LANES=2
run SCAPE_B200_STAGE_STEP=32 SCAPE_B200_LANES=3
SCAPE_B200_OVERLAP=0 SCAPE_B200_DBG=1 timeout 100 python bench.py --utrs 100 --per-file 1 --steps 1 --warmup 1 --no-cpu 2>&1 | grep -A2 "em run" | tail -12
