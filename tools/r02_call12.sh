#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_pins.py -x -q -m gpu > gpurun_out/r02_gputests_12.log 2>&1
tail -n 5 gpurun_out/r02_gputests_12.log | cut -c1-300
for c in 8 16 32 64; do
timeout 600 python bench.py --steps 3 --warmup 3 --chains $c --no-marginalised --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); r=d['roofline']; print('default chains $c ms_per_step',d['ms_per_step'],'state frac',r['frac'],'sweep_ms',r['sweep_ms'], 'value', d['value'])"
done
PGAS_STATE_SMALL=0 timeout 600 python bench.py --steps 3 --warmup 3 --chains 32 --no-marginalised --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); r=d['roofline']; print('big32 ms_per_step',d['ms_per_step'],'state frac',r['frac'],'sweep_ms',r['sweep_ms'], 'value', d['value'])"
PGAS_WEIGHTS_KERNEL=3 timeout 600 python bench.py --steps 3 --warmup 3 --chains 32 --no-marginalised --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); r=d['roofline']; print('lat32 ms_per_step',d['ms_per_step'],'state frac',r['frac'],'sweep_ms',r['sweep_ms'], 'value', d['value'])"
