#!/bin/bash
# chains-per-GPU sweep around the point where the state kernel switches to 256-thread CTAs (28 chains of N = 4096): the latency form of the
# resampling kernel is now the default only where its CTAs can share an SM with the state CTAs
mkdir -p gpurun_out
for c in 24 28 32; do
for wk in "" 3 1; do
PGAS_WEIGHTS_KERNEL=$wk timeout 300 python bench.py --chains $c --steps 3 --warmup 3 --no-marginalised --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); r=d['roofline']; print('[chains $c weights_kernel ${wk:-default}] ms_per_step %.2f sweep_ms %.2f value %.3e' % (d['ms_per_step'], r['sweep_ms'], d['value']))"
done; done | tee gpurun_out/r02_chains_28_32.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "split or sweep or run_chains" 2>&1 | tail -n 3 | cut -c1-300
