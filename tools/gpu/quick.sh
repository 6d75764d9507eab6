#!/bin/bash
# scratch: in-graph stage times, both dense schedules
mkdir -p gpurun_out
for v in 0 1; do
D3B_BEV_VARIANT=$v timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-nms-c5 > gpurun_out/bench_v$v.json 2> gpurun_out/bench_v$v.err; echo "bench v$v exit $?"; tail -2 gpurun_out/bench_v$v.err
done
python - <<PY
import json
for v in (0, 1):
    d=json.load(open("gpurun_out/bench_v%d.json" % v))
    print("v%d: value %.1f e2e %.1f ms %.4f | roofline %.1f TF/s frac %.4f launch_ms %.4f" % (v, d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["achieved"], d["roofline"]["frac"], d["roofline"]["launch_ms"]))
    print("   eager  ", {k: round(x,4) for k,x in d["stage_ms_per_step"].items()})
    print("   graph  ", {k: round(x,4) for k,x in (d["stage_ms_per_step_in_graph"] or {}).items()})
PY
