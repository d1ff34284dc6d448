#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python bench.py ) > gpurun_out/r02_bench_default_33.log 2> gpurun_out/r02_bench_default_33.err; tail -n 4 gpurun_out/r02_bench_default_33.err
python - <<'PY'
import json
for line in open('gpurun_out/r02_bench_default_33.log'):
    if line.startswith('{'):
        d=json.loads(line); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}); print(d['marginalised'].get('roofline')); print(d['roofline']['frac'], d['roofline']['sweep_frac'])
PY
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "split_and_fused or degenerate or dmma" 2>&1 | tail -2
