#!/bin/bash
# One GPU-box visit: parity tests, the benchmark line, the other configs, the ncu launch list.
mkdir -p gpurun_out
( timeout 420 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/pytest_gpu.log
timeout 300 python bench.py > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err
timeout 300 python bench.py --workload cfg3 --utrs 2000 --steps 1 --warmup 1 > gpurun_out/bench_cfg3_2k.json 2> gpurun_out/bench_cfg3_2k.err
timeout 200 python bench.py --workload cfg4 --utrs 2000 --steps 1 --warmup 1 > gpurun_out/bench_cfg4_2k.json 2> gpurun_out/bench_cfg4_2k.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches.csv \
  python bench.py --utrs 100 --per-file 1 --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_launch.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; tail -c 600 gpurun_out/bench_cfg2.json; tail -c 300 gpurun_out/bench_cfg3_2k.err
