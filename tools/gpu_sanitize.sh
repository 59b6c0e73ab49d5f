#!/bin/bash
# compute-sanitizer over the kernels added late in round 2 (small inputs)
set -u
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
run() { echo "--- $*"; timeout 400 "$@" 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Error|hazard" | head -8; }
run $S --tool memcheck python -m pytest tests/test_gpu_filter.py -q -x -k "store_rows_stream or split_by"
run $S --tool racecheck python -m pytest tests/test_gpu_filter.py -q -x -k "store_rows_stream"
run $S --tool memcheck python -m pytest tests/test_gpu_fullsize.py -q -x -k "metrics_of_a_whole"
run $S --tool racecheck python -m pytest tests/test_gpu_fullsize.py -q -x -k "metrics_of_a_whole"
run $S --tool memcheck python -m pytest tests/test_gpu_fast.py -q -x -k "fast_exact_counters_equal_oracle"
run $S --tool memcheck python -m pytest tests/test_gpu_step.py -q -x -k "overlapping"
