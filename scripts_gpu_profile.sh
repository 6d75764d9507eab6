#!/bin/bash
# launch list of one bench run (cold-cache, serialised: compare SHARES)
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
echo "ncu exit $?"
timeout 600 python -m pytest tests/test_spconv_gpu.py -m gpu -q --no-header --timeout 600 -p no:cacheprovider 2>&1 | tail -5
