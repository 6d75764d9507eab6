#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench3.json 2> gpurun_out/bench3.err; echo "bench exit $?"
tail -25 gpurun_out/bench3.err; head -c 3000 gpurun_out/bench3.json
timeout 600 python -m pytest tests/test_spconv_gpu.py -m gpu -q --no-header --timeout 600 -p no:cacheprovider -k "rulebook" 2>&1 | tail -3
