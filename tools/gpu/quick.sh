#!/bin/bash
mkdir -p gpurun_out
timeout 75 ncu --set full --clock-control none --import-source on -k regex:bev_conv16 -o gpurun_out/r2_prof_dense -f python scratch/ncu_dense_one.py > gpurun_out/ncu_dense.log 2>&1; echo "ncu exit $?"
ls -la gpurun_out/r2_prof_dense.ncu-rep
