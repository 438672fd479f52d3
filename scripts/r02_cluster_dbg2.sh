#!/bin/bash
mkdir -p gpurun_out
for C in 2 4 8; do
  SCAPE_B200_DBG=1 SCAPE_B200_OVERLAP=0 SCAPE_B200_CLUSTER=$C timeout 200 python bench.py --utrs 3000 --steps 1 --warmup 0 --no-cpu --no-cfg3 > gpurun_out/dbg2_c$C.json 2> gpurun_out/dbg2_c$C.err
  echo "== C=$C"; grep "cluster EM" gpurun_out/dbg2_c$C.err | tail -4 | cut -c1-330
done
