#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519"
timeout 300 $TR bench.py --gpus 8 --steps 2 --warmup 3 --no-cpu > gpurun_out/bench8_cfg2.json 2> gpurun_out/bench8_cfg2.err
timeout 200 $TR bench.py --gpus 8 --workload cfg3 --steps 1 --warmup 1 --no-cpu > gpurun_out/bench8_cfg3.json 2> gpurun_out/bench8_cfg3.err
nproc > gpurun_out/nproc.txt
for f in cfg2 cfg3; do python - $f <<'PY'
import json,sys
try:
    d=json.loads(open("gpurun_out/bench8_%s.json"%sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "n_gpus", d["n_gpus"], "value %.0f e2e %.0f scaling %s"%(d["value"], d["e2e"]["value"], d["scaling"]), d["phases_ms_per_step"])
except Exception as e:
    print(sys.argv[1], "FAILED", e, open("gpurun_out/bench8_%s.err"%sys.argv[1]).read()[-600:])
PY
done; cat gpurun_out/nproc.txt
