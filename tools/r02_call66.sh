#!/bin/bash
# new defaults at 28 / 32 / 36 / 40 chains per GPU, and the 64-thread state geometry (+ latency form) forced there
mkdir -p gpurun_out
run() { timeout 300 python bench.py --chains $1 --steps 3 --warmup 3 --no-marginalised --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); r=d['roofline']; print('[chains $1 $2] ms_per_step %.2f sweep_ms %.2f value %.3e' % (d['ms_per_step'], r['sweep_ms'], d['value']))"; }
for c in 28 32 40; do
run $c default
PGAS_STATE_SMALL=1 run $c state_small
done | tee gpurun_out/r02_chains_geometry.txt
