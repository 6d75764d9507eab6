#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_voxelize_gpu.py tests/test_pillars.py tests/test_e2e_gpu.py -m gpu -q --no-header --timeout 600 -p no:cacheprovider 2>&1 | tail -8
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench16.json 2> gpurun_out/bench16.err; echo "bench exit $?"
tail -3 gpurun_out/bench16.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench16.json"))
r=d["roofline"]; e=d["roofline_encoder"]
print("value %.1f e2e %.1f ms %.3f launches %s | bev3x3 %.1f TF/s frac %.3f launch_ms %.4f | enc %.1f GB/s ms %.3f | clocks %s"%(d["value"],d["e2e"]["value"],d["ms_per_step"],d["gpu_launches_per_step"],r["achieved"],r["frac"],r["launch_ms"],e["achieved"],e["kernel_ms_per_step"],d["clocks"]))
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 800 -c 400 --csv --log-file gpurun_out/launches_r1j.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/bench_under_ncu.log 2>&1
echo "ncu list exit $?"
python profiles/summarize_launches.py gpurun_out/launches_r1j.csv > gpurun_out/launches_r1j.md; head -16 gpurun_out/launches_r1j.md
