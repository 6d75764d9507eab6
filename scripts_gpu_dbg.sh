#!/bin/bash
mkdir -p gpurun_out
timeout 1100 compute-sanitizer --tool memcheck --print-limit 8 --launch-timeout 0 python -m pytest tests/test_e2e_gpu.py::test_cbgs_nuscenes_config_batch2 -m gpu -q --no-header --timeout 1000 -p no:cacheprovider -x > gpurun_out/dbg_san.log 2>&1
grep -n "Invalid\|Error\|at 0x\|by thread\|Address\|kernel" gpurun_out/dbg_san.log | head -40
tail -5 gpurun_out/dbg_san.log
