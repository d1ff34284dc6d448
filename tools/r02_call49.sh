#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --config 5 --steps 2 --warmup 2 --no-marginalised --no-cpu-baseline --no-strong 2>/dev/null | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); r=d['roofline']; print('[default] ms_per_step %.2f state frac %.4f sweep_ms %.2f sweep_frac %.3f' % (d['ms_per_step'], r['frac'], r['sweep_ms'], r['sweep_frac']))"
PGAS_SPLIT_TIMELINE=1 timeout 900 python bench.py --config 5 --steps 1 --warmup 1 --no-marginalised --no-cpu-baseline --no-strong 2>&1 | grep -A8 "^chunk" | head -9
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -n 3 | cut -c1-300
PGAS_WL_PPT_MIN=8 PGAS_WEIGHTS_KERNEL=3 timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "sweep or split or degenerate or full_size or run_chains or philox" 2>&1 | tail -n 3 | cut -c1-300
PGAS_WL_PPT_MIN=4 PGAS_WEIGHTS_KERNEL=3 timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "sweep or split or degenerate or full_size or run_chains or philox" 2>&1 | tail -n 3 | cut -c1-300
