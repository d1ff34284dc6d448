#!/bin/bash
# per-chunk timeline of the overlapped sweep at configs[3] (64 chains): when does state chunk c finish, when resampling chunk c?
mkdir -p gpurun_out
PGAS_SPLIT_TIMELINE=1 timeout 600 python bench.py --steps 1 --warmup 1 --no-marginalised --no-cpu-baseline 2>&1 | grep -A40 "^chunk" | tail -n 41 | tee gpurun_out/r02_timeline_cfg4.txt
PGAS_SPLIT_STATE_ROWS=8 timeout 600 python bench.py --steps 3 --warmup 3 --no-marginalised --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); r=d['roofline']; print('[state rows 8] ms_per_step %.2f sweep_ms %.2f' % (d['ms_per_step'], r['sweep_ms']))"
PGAS_SPLIT_ROWS=32 timeout 600 python bench.py --steps 3 --warmup 3 --no-marginalised --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); r=d['roofline']; print('[chunk rows 32] ms_per_step %.2f sweep_ms %.2f' % (d['ms_per_step'], r['sweep_ms']))"
PGAS_SPLIT_ROWS=128 timeout 600 python bench.py --steps 3 --warmup 3 --no-marginalised --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); r=d['roofline']; print('[chunk rows 128] ms_per_step %.2f sweep_ms %.2f' % (d['ms_per_step'], r['sweep_ms']))"
