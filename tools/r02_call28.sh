#!/bin/bash
mkdir -p gpurun_out
for e in "A=1" "PGAS_SPLIT_PRE_C=2" "PGAS_SPLIT_PRE_C=8" "PGAS_SPLIT_PRE_C=16" "PGAS_SPLIT_STATE_ROWS=8" "PGAS_SPLIT_STATE_ROWS=32"; do
env $e timeout 900 python bench.py --config 5 --steps 2 --warmup 2 --no-cpu-baseline --no-strong 2>/dev/null | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); r=d['roofline']; print('[$e] cfg5 ms_per_step %.1f value %.3e state frac %.3f sweep_ms %.1f sweep_frac %.3f' % (d['ms_per_step'],d['value'],r['frac'],r['sweep_ms'],r['sweep_frac']))"
done
