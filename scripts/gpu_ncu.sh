#!/bin/bash
# One ncu --set full capture per hot kernel on a single cfg-2 wave (100 UTRs x 500 reads, 100 streams).
# The first pass (warm-up) launches 1 table + 2 tensor + ~103 EM kernels; captures are taken in the second pass.
# The .ncu-rep files are ~35 MB each and gpurun_out/ carries 64 MiB back: export the raw and source
# pages as CSV on the box and keep only one report.
mkdir -p gpurun_out
CMD="python bench.py --utrs 100 --per-file 1 --steps 1 --warmup 1 --no-cpu"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
cap() {  # cap <name> <kernel regex> <launch skip>
  timeout 250 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -f -o /tmp/prof_$1 $CMD > gpurun_out/ncu_$1.log 2>&1
  ncu -i /tmp/prof_$1.ncu-rep --page raw --csv > gpurun_out/ncu_$1_raw.csv 2>/dev/null
  ncu -i /tmp/prof_$1.ncu-rep --page source --csv > gpurun_out/ncu_$1_source.csv 2>/dev/null
  ncu -i /tmp/prof_$1.ncu-rep --page details > gpurun_out/ncu_$1_details.txt 2>/dev/null
}
# warp E step: 24 launches per pass -> step 1 of the second pass; CTA E step: 27 per pass -> step 24; scan: 50 per pass -> step 1
cap estep_warp em_estep_warp_kernel 25
cap estep_cta em_estep_kernel 27
cap scan em_scan_kernel 51
gzip -f gpurun_out/ncu_*_source.csv
ls -la gpurun_out/ | tail -15
