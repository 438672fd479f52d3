#!/bin/bash
# GPU-box visit: smoke, parity tests, benchmark line (cfg-2 + cfg3 block + cpu_baseline leg), reference arm.
# Usage: scripts/r02_run.sh <tag>
tag=${1:-r02}
mkdir -p gpurun_out
timeout 180 python __graft_entry__.py --smoke > gpurun_out/${tag}_smoke.log 2>&1; rc=$?; echo "smoke rc=$rc"; tail -1 gpurun_out/${tag}_smoke.log
if [ $rc -ne 0 ]; then tail -20 gpurun_out/${tag}_smoke.log; exit 1; fi
(timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -70) > gpurun_out/${tag}_pytest.log; tail -30 gpurun_out/${tag}_pytest.log | cut -c1-330
t0=$(date +%s); timeout 900 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench.py wall: $(( $(date +%s) - t0 )) s"
timeout 600 python bench.py --impl reference > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${tag}_bench.json").read().strip().splitlines()[-1])
    print("cfg2 value %.0f e2e %.0f roof %.3f launches %d"%(d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["gpu_launches"]), d.get("cpu_baseline"), d.get("parity_check"))
    print({k: round(v) for k, v in d["phases_alone_ms"].items()})
    c=d.get("cfg3")
    if c: print("cfg3 value %.0f e2e %.0f roof %.3f waves %d"%(c["value"], c["e2e"]["value"], c["roofline"]["frac"], c["waves_per_step"]), {k: round(v) for k, v in c["phases_alone_ms"].items()})
    r=json.loads(open("gpurun_out/${tag}_bench_reference.json").read().strip().splitlines()[-1])
    print("reference arm value %.2f"%r["value"], r["cpu_baseline"]["cores"], "cores; same workload string:", r["config"]["workload"]==d["config"]["workload"])
except Exception as e:
    print("bench FAILED", e); print(open("gpurun_out/${tag}_bench.err").read()[-1500:])
PY
