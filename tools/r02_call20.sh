#!/bin/bash
mkdir -p gpurun_out
timeout 1700 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "plugin or suffstats_and_draw or run_chains_matches or basis" > gpurun_out/r02_gputests_20.log 2>&1
tail -n 15 gpurun_out/r02_gputests_20.log | cut -c1-300
timeout 600 python bench.py --steps 3 --warmup 3 --no-marginalised --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); r=d['roofline']; print('cfg4 ms_per_step',d['ms_per_step'],'state frac',r['frac'],'sweep_ms',r['sweep_ms'], 'value', d['value'], 'share', d.get('split_8gpu_share'))"
