#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_filter.py -q --tb=short -x -k "store_rows" 2>&1 | tail -3
(timeout 120 python tools/run_store.py; BLP_STORE_STREAM=0 timeout 120 python tools/run_store.py) 2>&1 | tee gpurun_out/store_rows.txt
