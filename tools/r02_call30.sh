#!/bin/bash
mkdir -p gpurun_out
PGAS_STATE_MMA=1 PGAS_STATE_SMALL=0 timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "sweep_parity_injected or philox_stream or run_chains_matches or full_size" > gpurun_out/r02_gputests_30.log 2>&1
tail -n 8 gpurun_out/r02_gputests_30.log | cut -c1-250
for e in "PGAS_STATE_MMA=0" "PGAS_STATE_MMA=1"; do
env $e timeout 600 python tools/state_probe.py 64 401 4 2>&1 | tail -n 1 | cut -c1-330
env $e timeout 600 python tools/state_probe.py 16 201 5 2>&1 | tail -n 1 | cut -c1-330
env $e timeout 600 python bench.py --steps 3 --warmup 3 --no-marginalised --no-cpu-baseline --no-strong 2>/dev/null | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); r=d['roofline']; print('[$e] ms_per_step %.2f state frac %.4f sweep_ms %.2f value %.4e' % (d['ms_per_step'], r['frac'], r['sweep_ms'], d['value']))"
done
