#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1100 -c 500 --csv --log-file gpurun_out/launches_r1c.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/bench_under_ncu.log 2>&1
echo "ncu exit $?"
