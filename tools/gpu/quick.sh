#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_spconv_gpu.py tests/test_e2e_gpu.py -m gpu -q --no-header --timeout 600 -p no:cacheprovider -x 2>&1 | tail -4
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench17.json 2> gpurun_out/bench17.err; echo "bench exit $?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench17.json"))
r=d["roofline"]; e=d["roofline_encoder"]
print("value %.1f e2e %.1f ms %.3f | bev3x3 %.1f TF/s frac %.3f launch_ms %.4f | enc ms %.3f"%(d["value"],d["e2e"]["value"],d["ms_per_step"],r["achieved"],r["frac"],r["launch_ms"],e["kernel_ms_per_step"]))
PY
timeout 200 python scratch/trace_run.py 2>&1 | grep "MMA issue"
