#!/bin/bash
mkdir -p gpurun_out
timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 --steps 30 --warmup 5 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err; echo "bench N=4 exit $?"
python - <<PY
import json
raw=open("gpurun_out/bench_n4.json").read()
print("lines", raw.count("\n"))
d=json.loads(raw)
print("N=%d value %.1f e2e %.1f ms %.3f"%(d["n_gpus"],d["value"],d["e2e"]["value"],d["ms_per_step"]))
PY
