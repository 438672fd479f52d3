#!/bin/bash
mkdir -p gpurun_out
for U in 2500 5000 10000; do
  SCAPE_B200_EM=bsp SCAPE_B200_DBG=1 SCAPE_B200_OVERLAP=0 timeout 200 python bench.py --utrs $U --steps 1 --warmup 0 --no-cpu --no-cfg3 > gpurun_out/bsp_u$U.json 2> gpurun_out/bsp_u$U.err
  echo "== UTRS=$U"; grep -A2 "em run: small=[0-9]* big=0 refs=[1-9]" gpurun_out/bsp_u$U.err | sed -n 7,9p | cut -c1-700
done
