#!/bin/bash
set -u
mkdir -p gpurun_out
for n in 600000 2400000; do
  timeout 600 python bench.py --steps 3 --warmup 3 --no-legs --no-cpu-baseline --wd-entities $n 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); w=d['wikidata5m_scale_sweep']
print('wd entities', w['entities'], 'pass2 ms/pass', w['table_pass_2']['ms_per_pass'], 'hbm_frac', w['table_pass_2']['hbm_frac'], '| value', d['value'], 'e2e', d['e2e']['value'])"
done
