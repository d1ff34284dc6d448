#!/bin/bash
mkdir -p gpurun_out
for env in "" "PGAS_SPLIT_SERIAL=1" "PGAS_SPLIT_ROWS=16" "PGAS_SPLIT_ROWS=32" "PGAS_SPLIT_STATE_ROWS=64" "PGAS_SPLIT_ONE_GROUP=1"; do
env $env timeout 600 python bench.py --steps 3 --warmup 3 --chains 8 --no-marginalised --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); r=d['roofline']; print('[$env] chains 8 ms_per_step',d['ms_per_step'],'state frac',r['frac'],'sweep_ms',r['sweep_ms'], 'launches', d['gpu_launches'])"
done
