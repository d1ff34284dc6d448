#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -n 3 | cut -c1-250
timeout 600 python tools/state_probe.py 64 401 4 2>&1 | tail -n 1 | cut -c1-330
timeout 600 python tools/state_probe.py 16 201 5 2>&1 | tail -n 1 | cut -c1-330
timeout 600 python bench.py --steps 3 --warmup 3 --no-marginalised --no-cpu-baseline --no-strong 2>/dev/null | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); r=d['roofline']; print('ms_per_step %.2f state frac %.4f sweep_ms %.2f value %.4e' % (d['ms_per_step'], r['frac'], r['sweep_ms'], d['value']))"
