#!/bin/bash
# scratch: sparse gather (skip cells that stay zero) -- parity + in-graph stage times, all three configs
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_conv16_gpu.py tests/test_spconv_gpu.py -x -q -m gpu --no-header -p no:cacheprovider 2>&1 | tail -4
for c in second pillars cbgs; do
timeout 300 python bench.py --config $c --steps 30 --warmup 5 --no-cpu-baseline --no-nms-c5 > gpurun_out/bench_q_$c.json 2> gpurun_out/bench_q_$c.err; echo "bench $c exit $?"; tail -1 gpurun_out/bench_q_$c.err
done
python - <<PY
import json
for c in ("second", "pillars", "cbgs"):
    try:
        d=json.load(open("gpurun_out/bench_q_%s.json" % c))
    except Exception as ex:
        print(c, "no json", ex); continue
    print("%s: value %.1f e2e %.1f ms %.4f | %s | %.1f TF/s frac %.4f launch_ms %.4f" % (c, d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["kernel"][:60], d["roofline"]["achieved"], d["roofline"]["frac"], d["roofline"]["launch_ms"]))
    print("   eager  ", {k: round(x,4) for k,x in d["stage_ms_per_step"].items()})
    print("   graph  ", {k: round(x,4) for k,x in (d["stage_ms_per_step_in_graph"] or {}).items()})
PY
