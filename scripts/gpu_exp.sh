#!/bin/bash
# development aid: A/B runs of env knobs on one GPU box; summary lines in gpurun_out/exp.txt
mkdir -p gpurun_out
: > gpurun_out/exp.txt
run() {  # run "<env assignments>" <bench args...>
  local envs="$1"; shift
  env $envs timeout 300 python bench.py --no-cpu "$@" > gpurun_out/sw.log 2> gpurun_out/sw.err
  python - "$envs | $*" <<PY >> gpurun_out/exp.txt
import json,sys
try:
    d=json.loads(open("gpurun_out/sw.log").read().strip().splitlines()[-1])
    p=d["phases_ms_per_step"]; a=d["phases_alone_ms"]
    print(sys.argv[1], "| value %.0f e2e %.0f busy %.0f total %.0f | em %.0f estep %.0f scan %.0f tensor %.0f table %.0f rng %.0f prep %.0f | alone: estep %.0f scan %.0f | roof %.3f" % (d["value"], d["e2e"]["value"], p["device_busy_ms"], p["total_ms"], p["em_ms"], p["estep_ms"], p["scan_ms"], p["tensor_ms"], p["table_ms"], p["host_rng_ms"], p["host_prep_ms"], a["estep_ms"], a["scan_ms"], d["roofline"]["frac"]))
except Exception as e:
    print(sys.argv[1], "FAILED", e, open("gpurun_out/sw.err").read()[-400:])
PY
}
( timeout 300 python -m pytest tests -m gpu -x -q -k "chain_traces or goldens or ragged or many_streams" 2>&1 | tail -5 ) > gpurun_out/pytest_gpu.log
run "A=1" --steps 2 --warmup 2
run "SCAPE_B200_WARP_MAXN=512" --steps 2 --warmup 2
C3="--workload cfg3 --utrs 6000 --steps 1 --warmup 1"
run "A=1" $C3
run "SCAPE_B200_SCAN_SPLIT=0" $C3
run "SCAPE_B200_WARP_MAXN=512" $C3
run "SCAPE_B200_WARP_MAXN=256" $C3
run "SCAPE_B200_WARP_STEPS=0" $C3
run "SCAPE_B200_LANES=2" $C3
run "SCAPE_B200_LANES=4" $C3
SCAPE_B200_OVERLAP=0 SCAPE_B200_DBG=1 timeout 200 python bench.py --workload cfg3 --utrs 2000 --steps 1 --warmup 0 --no-cpu > /dev/null 2> gpurun_out/dbg_cfg3.txt
SCAPE_B200_OVERLAP=0 SCAPE_B200_DBG=1 timeout 100 python bench.py --utrs 100 --per-file 1 --steps 1 --warmup 1 --no-cpu 2>&1 | grep -A2 "em run" | tail -12 > gpurun_out/dbg_cfg2.txt
cat gpurun_out/pytest_gpu.log gpurun_out/exp.txt
