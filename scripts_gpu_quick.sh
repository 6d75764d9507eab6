#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ingest.py tests/test_voxelize_gpu.py -m gpu -q --no-header --timeout 600 -p no:cacheprovider 2>&1 | tail -25
