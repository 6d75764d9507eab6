#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_spconv_gpu.py tests/test_voxelize_gpu.py -m gpu -q --no-header --timeout 300 -p no:cacheprovider --maxfail=60 > gpurun_out/pytest_tc.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_tc.log
grep -E "passed|failed" gpurun_out/pytest_tc.log | tail -3
grep -E "^FAILED" gpurun_out/pytest_tc.log | head -40
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tc.json 2> gpurun_out/bench_tc.err; echo "bench exit $?"
cat gpurun_out/bench_tc.json | head -c 2500; tail -5 gpurun_out/bench_tc.err
