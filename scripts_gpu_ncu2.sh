#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spconv_tc_kernel -s 44 -c 22 -o gpurun_out/prof_spconv_tc_v3 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"; ls -la gpurun_out/prof_spconv_tc_v3.ncu-rep
