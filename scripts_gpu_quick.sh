#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_spconv_gpu.py tests/test_e2e_gpu.py -m gpu -q --no-header --timeout 600 -p no:cacheprovider -x 2>&1 | tail -15
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench13.json 2> gpurun_out/bench13.err; echo "bench exit $?"
tail -3 gpurun_out/bench13.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench13.json"))
print("value %.1f e2e %.1f ms %.3f launches %s enc_ms %.3f bev_ms %s"%(d["value"],d["e2e"]["value"],d["ms_per_step"],d["gpu_launches_per_step"],d["roofline"]["kernel_ms_per_step"], d.get("roofline_bev",{}).get("kernel_ms_per_step")))
PY
timeout 300 python scratch/trace_run.py 2>&1 | tail -150 > gpurun_out/trace2.log
