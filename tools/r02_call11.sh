#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "split_and_fused_sweeps_agree" > gpurun_out/r02_gputests_11.log 2>&1
tail -n 25 gpurun_out/r02_gputests_11.log | cut -c1-300
for c in 8 64; do
PGAS_WEIGHTS_KERNEL=3 timeout 600 python bench.py --steps 3 --warmup 3 --chains $c --no-marginalised --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); r=d['roofline']; print('lat   chains $c ms_per_step',d['ms_per_step'],'state frac',r['frac'],'sweep_ms',r['sweep_ms'])"
PGAS_WEIGHTS_KERNEL=1 timeout 600 python bench.py --steps 3 --warmup 3 --chains $c --no-marginalised --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); r=d['roofline']; print('1-cta chains $c ms_per_step',d['ms_per_step'],'state frac',r['frac'],'sweep_ms',r['sweep_ms'])"
done
