#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu --no-header --timeout 600 -p no:cacheprovider 2>&1 | tail -6
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench exit $?"
tail -2 gpurun_out/bench_final.err
python - <<PY
import json
raw=open("gpurun_out/bench_final.json").read()
assert raw.count("\n")==1, raw[:200]
d=json.loads(raw)
print({k:d[k] for k in ("value","ms_per_step","steps","gpu_launches","clocks")}, d["e2e"], d["cpu_baseline"])
print(d["roofline"]["achieved"], d["roofline"]["frac"], d["roofline"]["traffic"], d["roofline_encoder"]["achieved"])
PY
