#!/bin/bash
# One GPU-box visit for the record: parity tests, the benchmark line + reference arm, the other
# BASELINE.json configs, the ncu launch list and the scan's DRAM traffic per launch.
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/pytest_gpu.log
timeout 400 python bench.py > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err
timeout 300 python bench.py --impl reference > gpurun_out/bench_reference_arm.json 2> gpurun_out/bench_reference_arm.err
timeout 400 python bench.py --workload cfg3 --steps 1 --warmup 1 > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err
timeout 300 python bench.py --workload cfg4 --steps 1 --warmup 1 > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err
WAVE="python bench.py --utrs 100 --per-file 1 --steps 1 --warmup 1 --no-cpu"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv $WAVE > gpurun_out/ncu_launch.log 2>&1
timeout 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:em_scan_kernel -c 150 --csv --log-file gpurun_out/scan_dram.csv $WAVE > gpurun_out/ncu_scan_dram.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; for f in cfg2 cfg3 cfg4; do python - $f <<'PY'
import json,sys
try:
    d=json.loads(open("gpurun_out/bench_%s.json"%sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value %.0f e2e %.0f roof %.3f"%(d["value"], d["e2e"]["value"], d["roofline"]["frac"]), d.get("cpu_baseline"))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done; tail -c 300 gpurun_out/bench_reference_arm.json
