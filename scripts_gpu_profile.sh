#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 800 -c 400 --csv --log-file gpurun_out/launches_r1d.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/bench_under_ncu.log 2>&1
echo "ncu list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"spconv_pairs_kernel|spconv_tc_kernel" -s 44 -c 22 -o gpurun_out/prof_conv_r1d -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"
