#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_pillars.py tests/test_e2e_gpu.py tests/test_nms_gpu.py -m gpu -q --no-header --timeout 600 -p no:cacheprovider 2>&1 | tail -25
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench7.json 2> gpurun_out/bench7.err; echo "bench exit $?"
tail -3 gpurun_out/bench7.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench7.json"))
print("value %.1f e2e %.1f ms %.3f launches %s enc_ms %.3f bev_ms %s"%(d["value"],d["e2e"]["value"],d["ms_per_step"],d["gpu_launches_per_step"],d["roofline"]["kernel_ms_per_step"], d.get("roofline_bev",{}).get("kernel_ms_per_step")))
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 800 -c 400 --csv --log-file gpurun_out/launches_r1e.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/bench_under_ncu.log 2>&1
echo "ncu list exit $?"
