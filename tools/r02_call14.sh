#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "split_and_fused_sweeps_agree" > gpurun_out/r02_gputests_14.log 2>&1
tail -n 12 gpurun_out/r02_gputests_14.log | cut -c1-300
timeout 600 python tools/ticks_bench.py 2 8 > gpurun_out/r02_ticks_14.log 2>&1
head -12 gpurun_out/r02_ticks_14.log
for c in 8 16; do
timeout 600 python bench.py --steps 3 --warmup 3 --chains $c --no-marginalised --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); r=d['roofline']; print('default chains $c ms_per_step',d['ms_per_step'],'state frac',r['frac'],'sweep_ms',r['sweep_ms'], 'value', d['value'])"
done
