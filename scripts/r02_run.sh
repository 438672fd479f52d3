#!/bin/bash
# GPU-box visit: smoke (with a short timeout: a wrong cluster barrier would hang), memcheck of the smoke,
# parity tests, benchmark line.  Usage: scripts/r02_run.sh <tag>
tag=${1:-r02}
mkdir -p gpurun_out
timeout 180 python __graft_entry__.py --smoke > gpurun_out/${tag}_smoke.log 2>&1; rc=$?; echo "smoke rc=$rc"; tail -3 gpurun_out/${tag}_smoke.log
if [ $rc -ne 0 ]; then exit 1; fi
if [ -n "$MEMCHECK" ]; then
  timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python __graft_entry__.py --smoke > gpurun_out/${tag}_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|Invalid|error" gpurun_out/${tag}_memcheck.log | head -20
fi
(timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -70) > gpurun_out/${tag}_pytest.log; tail -45 gpurun_out/${tag}_pytest.log | cut -c1-400
timeout 400 python bench.py --no-cpu > gpurun_out/${tag}_bench_cfg2.json 2> gpurun_out/${tag}_bench_cfg2.err; tail -2 gpurun_out/${tag}_bench_cfg2.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${tag}_bench_cfg2.json").read().strip().splitlines()[-1])
    print("cfg2 value %.0f e2e %.0f roof %.3f"%(d["value"], d["e2e"]["value"], d["roofline"]["frac"]))
    print(d["phases_ms_per_step"]); print(d["phases_alone_ms"])
except Exception as e:
    print("bench FAILED", e)
PY
