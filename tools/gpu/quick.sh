#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --dist uniform --steps 20 --warmup 4 --no-cpu-baseline > gpurun_out/bench_uniform.json 2>gpurun_out/bench_uniform.err; echo exit $?
python - <<PY
import json
d=json.load(open("gpurun_out/bench_uniform.json")); e=d["roofline_encoder"]
print("uniform: value %.1f e2e %.1f ms %.3f | enc %.1f GB/s, %.3f ms, bytes %.1f MB, %.1f GFLOP" % (d["value"], d["e2e"]["value"], d["ms_per_step"], e["achieved"], e["kernel_ms_per_step"], e["algorithmic_bytes_per_step"]/1e6, e["flops_per_step"]/1e9))
PY
