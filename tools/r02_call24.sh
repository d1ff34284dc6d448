#!/bin/bash
mkdir -p gpurun_out
timeout 1700 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "split_and_fused or degenerate" > gpurun_out/r02_gputests_24.log 2>&1
tail -n 6 gpurun_out/r02_gputests_24.log | cut -c1-300
for env in "PGAS_WEIGHTS_KERNEL=3" "A=1"; do
env $env timeout 900 python bench.py --config 5 --steps 2 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); r=d['roofline']; print('[$env] cfg5 ms_per_step',d['ms_per_step'],'value',d['value'],'state frac',r['frac'],'sweep_ms',r['sweep_ms'],'sweep_frac',r['sweep_frac'])"
done
