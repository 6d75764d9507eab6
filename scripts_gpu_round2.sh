#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --no-header --timeout 600 -p no:cacheprovider --maxfail=40 > gpurun_out/pytest_gpu2.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu2.log
grep -E "passed|failed" gpurun_out/pytest_gpu2.log | tail -3
grep -E "^FAILED|^ERROR" gpurun_out/pytest_gpu2.log | head -40
for flags in "" "--no-graph" "--no-fused-bev" "--no-graph --no-fused-bev"; do
  timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline $flags > gpurun_out/bench2_$(echo $flags | tr -d ' -').json 2> gpurun_out/bench2.err; echo "bench [$flags] exit $?"
  python - <<PY
import json
d=json.load(open("gpurun_out/bench2_$(echo $flags | tr -d ' -').json"))
print("value %.1f e2e %.1f ms %.3f launches/step %s enc_ms %.3f bev %s"%(d["value"],d["e2e"]["value"],d["ms_per_step"],d["gpu_launches_per_step"],d["roofline"]["kernel_ms_per_step"], d.get("roofline_bev",{}).get("kernel_ms_per_step")))
PY
  tail -2 gpurun_out/bench2.err
done
