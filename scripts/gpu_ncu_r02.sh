#!/bin/bash
# Round-2 ncu evidence on a single cfg-2 wave (100 UTRs x 500 reads, 100 streams = one wave, split in two halves):
# the launch list of the timed pass, and one --set full capture per kernel of the likelihood phase, the scan and the E step.
mkdir -p gpurun_out
# (one lane, so that the launch order is the same in every run: parts of a split wave are enqueued one after the other)
export SCAPE_B200_LANES=1
CMD="python bench.py --utrs 100 --per-file 1 --steps 1 --warmup 1 --no-cpu --no-cfg3 --no-files"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
cap() {  # cap <name> <kernel regex> <launch skip>
  timeout 250 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -f -o /tmp/prof_$1 $CMD > gpurun_out/ncu_$1.log 2>&1
  ncu -i /tmp/prof_$1.ncu-rep --page raw --csv > gpurun_out/r02_ncu_$1_raw.csv 2>/dev/null
  ncu -i /tmp/prof_$1.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/r02_ncu_$1_source.csv.gz
}
# one pass = 1 table + 1 tensor_interior launch (all alpha rows since round 2c); the second pass (timed) is captured
cap table table_kernel 1
cap tensor_interior tensor_interior_kernel 1
# two halves x 50 scans / 32 warp E steps per pass: second pass, step 1 of the first half
cap scan em_scan_kernel 101
cap estep_warp em_estep_warp_kernel 65
ls -la gpurun_out/ | grep r02_ | head -20
