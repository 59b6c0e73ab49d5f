#!/bin/bash
set -u
mkdir -p gpurun_out
BLP_SWEEP_CFG=5 timeout 600 python -m pytest tests/test_gpu_step.py tests/test_gpu_eval.py tests/test_gpu_fullsize.py -q --tb=short 2>&1 | tail -3
for cfg in 4 5; do for m in transe distmult; do
  BLP_SWEEP_CFG=$cfg timeout 120 python tools/run_step.py $m 1024 14541 30 2>&1 | tail -1 | cut -c1-140 | sed "s/^/cfg$cfg: /"
done; done
for cfg in 2 3 5; do
  BLP_SWEEP_CFG=$cfg timeout 120 python tools/run_step.py transe 64 14541 200 2>&1 | tail -1 | cut -c1-140 | sed "s/^/cfg$cfg: /"
done
for cfg in 4 5; do
BLP_SWEEP_CFG=$cfg timeout 300 python bench.py --steps 5 --warmup 3 --no-legs --no-cpu-baseline --no-extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('cfg$cfg bench value', d['value'], 'e2e', d['e2e']['value'], 'kernel_ms', r['kernel_ms'], 'frac', r['frac'])"
done
