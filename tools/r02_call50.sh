#!/bin/bash
for e in "PGAS_X=1" "PGAS_WEIGHTS_KERNEL=0"; do
env $e timeout 600 python bench.py --config 5 --steps 2 --warmup 2 --no-marginalised --no-cpu-baseline --no-strong 2>/dev/null | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); r=d['roofline']; print('[$e] ms_per_step %.2f state frac %.4f sweep_ms %.2f sweep_frac %.3f' % (d['ms_per_step'], r['frac'], r['sweep_ms'], r['sweep_frac']))"
done
PGAS_SPLIT_TIMELINE=1 timeout 900 python bench.py --config 5 --steps 1 --warmup 1 --no-marginalised --no-cpu-baseline --no-strong 2>&1 | grep -A6 "^chunk" | head -7
