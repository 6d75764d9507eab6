#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -4
for N in 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench N=$N exit $?"
tail -4 gpurun_out/bench_n$N.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_n$N.json"))
print("N=%d value %.1f e2e %.1f ms %.3f"%(d["n_gpus"],d["value"],d["e2e"]["value"],d["ms_per_step"]))
PY
done
timeout 300 python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline | python -c "import json,sys; d=json.load(sys.stdin); print('N=1 value %.1f e2e %.1f'%(d['value'],d['e2e']['value']))"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --impl reference --steps 3 --warmup 1 2>&1 | tail -2 | cut -c1-400
