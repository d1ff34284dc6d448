#!/bin/bash
mkdir -p gpurun_out
PGAS_SPLIT_TIMELINE=1 timeout 900 python bench.py --config 5 --steps 1 --warmup 1 --no-marginalised --no-cpu-baseline --no-strong 2>&1 | grep -A14 "^chunk" | head -48
for e in "PGAS_SPLIT_SERIAL=1" "PGAS_SPLIT_STATE_ROWS=8" "PGAS_SPLIT_STATE_ROWS=32" "PGAS_SPLIT_ROWS=32" "PGAS_WEIGHTS_KERNEL=3"; do
env $e timeout 600 python bench.py --config 5 --steps 2 --warmup 2 --no-marginalised --no-cpu-baseline --no-strong 2>/dev/null | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); r=d['roofline']; print('[$e] ms_per_step %.2f state frac %.4f sweep_ms %.2f' % (d['ms_per_step'], r['frac'], r['sweep_ms']))"
done
