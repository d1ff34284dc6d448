#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -n 3 | cut -c1-250
timeout 600 python bench.py --steps 3 --warmup 3 --no-marginalised --no-cpu-baseline --no-strong 2>/dev/null | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); r=d['roofline']; print('ms_per_step %.2f state frac %.4f sweep_ms %.2f value %.4e' % (d['ms_per_step'], r['frac'], r['sweep_ms'], d['value'])); print(json.dumps(d.get('roofline_kernels'))[:1500])"
