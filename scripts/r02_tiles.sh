#!/bin/bash
mkdir -p gpurun_out
timeout 180 python __graft_entry__.py --smoke 2>&1 | tail -1
(timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -8) | cut -c1-900
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu --no-cfg3 --steps 2 --warmup 1 > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$name.json").read().strip().splitlines()[-1])
    print("$name: value %.0f e2e %.0f roof %.3f"%(d["value"], d["e2e"]["value"], d["roofline"]["frac"]), {k: round(v) for k, v in d["phases_alone_ms"].items()})
except Exception as e:
    print("$name FAILED", e); print(open("gpurun_out/bench_$name.err").read()[-800:])
PY
}
run bsp_old SCAPE_B200_EM=bsp SCAPE_B200_SCAN_TILES=0
run bsp_tiles SCAPE_B200_EM=bsp
run tail24 SCAPE_B200_TAIL_STEP=24
