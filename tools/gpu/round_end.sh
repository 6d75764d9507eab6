#!/bin/bash
# Round-end evidence run (one gpurun call): development trace, full -m gpu suite, smoke(), the three bench configs,
# launch list and ncu --set full of the convolution kernels.  Ordered by priority; every step has its own timeout.
mkdir -p gpurun_out
timeout 120 python scratch/chain_trace.py > gpurun_out/chain_trace2.log 2>&1; echo "chain trace exit $?"
timeout 560 python -m pytest tests -x -q -m gpu --no-header --timeout 300 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu.log
timeout 150 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 400 python bench.py > gpurun_out/r2_bench_second.json 2> gpurun_out/bench_second.err; echo "bench second exit $?"; tail -1 gpurun_out/bench_second.err
python - <<PY
import json
raw=open("gpurun_out/r2_bench_second.json").read()
assert raw.count("\n")==1, raw[:200]
d=json.loads(raw)
print({k:d[k] for k in ("value","ms_per_step","steps","gpu_launches","clocks")}, d["e2e"], {k: d["cpu_baseline"].get(k) for k in ("value","cores","kind")})
print(d["roofline"]["kernel"][:50], d["roofline"]["achieved"], d["roofline"]["frac"], d["roofline"]["launch_ms"], "| enc", d["roofline_encoder"]["achieved"], d["roofline_encoder"]["kernel_ms_per_step"])
print("graph", {k: round(x,4) for k,x in (d["stage_ms_per_step_in_graph"] or {}).items()})
PY
for c in pillars cbgs; do
timeout 300 python bench.py --config $c --no-nms-c5 > gpurun_out/r2_bench_$c.json 2> gpurun_out/bench_$c.err; echo "bench $c exit $?"
done
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launch_list.csv python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline --no-nms-c5 > gpurun_out/ncu_list.log 2>&1; echo "ncu list exit $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"spconv_os16_kernel|bev_conv16" -s 44 -c 22 -o gpurun_out/r2_prof_conv -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-nms-c5 --no-graph > gpurun_out/ncu_full.log 2>&1; echo "ncu full exit $?"
ls -la gpurun_out/*.ncu-rep 2>/dev/null
