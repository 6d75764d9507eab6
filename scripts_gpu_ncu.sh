#!/bin/bash
# (1) launch list of the bench (shares), (2) ncu --set full of the dominant sparse-conv launches
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 450 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
echo "ncu list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spconv_tc_kernel -s 28 -c 14 -o gpurun_out/prof_spconv_tc -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"vox_|rb_|nms_|sparse_to_dense" -s 60 -c 30 -o gpurun_out/prof_misc -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full2.log 2>&1
echo "ncu full2 exit $?"
ls -la gpurun_out/*.ncu-rep
