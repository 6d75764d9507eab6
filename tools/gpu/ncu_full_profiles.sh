#!/bin/bash
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"spconv_pairs_kernel|spconv_tc_kernel" -s 44 -c 22 -o gpurun_out/prof_conv_r1_final -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"; ls -la gpurun_out/prof_conv_r1_final.ncu-rep
timeout 600 ncu --set full --clock-control none -k regex:"nms_mask_kernel|nms_sweep_kernel|topk_|vox_|rb_neighbours|rb_mark|rb_emit|decode_selected|head_scores|finalize" -s 120 -c 40 -o gpurun_out/prof_misc_r1_final -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_full2.log 2>&1
echo "ncu misc exit $?"; ls -la gpurun_out/prof_misc_r1_final.ncu-rep
