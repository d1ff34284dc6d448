#!/bin/bash
mkdir -p gpurun_out
for c in 32 48 64; do for env in "PGAS_WEIGHTS_KERNEL=3" "PGAS_WEIGHTS_KERNEL=1" "PGAS_WEIGHTS_KERNEL=3 PGAS_STATE_SMALL=1"; do
env $env timeout 600 python bench.py --steps 3 --warmup 3 --chains $c --no-marginalised --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); r=d['roofline']; print('[$env] chains $c ms_per_step',d['ms_per_step'],'state frac',r['frac'],'sweep_ms',r['sweep_ms'], 'value', d['value'])"
done; done
