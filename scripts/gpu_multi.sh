#!/bin/bash
# torchrun lines on N GPUs of one box: cfg-2 (weak scaling) and cfg-3 (strong scaling, 20k UTRs).
# usage (under gpurun --gpus N): bash scripts/gpu_multi.sh N
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 400 $TR bench.py --gpus $N --steps 2 --warmup 3 --no-cpu > gpurun_out/bench${N}_cfg2.json 2> gpurun_out/bench${N}_cfg2.err
timeout 400 $TR bench.py --gpus $N --workload cfg3 --steps 1 --warmup 1 --no-cpu > gpurun_out/bench${N}_cfg3.json 2> gpurun_out/bench${N}_cfg3.err
nproc > gpurun_out/nproc.txt
for f in cfg2 cfg3; do python - $N $f <<'PY'
import json,sys
n,f=sys.argv[1],sys.argv[2]
try:
    d=json.loads(open("gpurun_out/bench%s_%s.json"%(n,f)).read().strip().splitlines()[-1])
    print(f, "n_gpus", d["n_gpus"], "value %.0f e2e %.0f scaling %s"%(d["value"], d["e2e"]["value"], d["scaling"]))
except Exception as e:
    print(f, "FAILED", e, open("gpurun_out/bench%s_%s.err"%(n,f)).read()[-600:])
PY
done
