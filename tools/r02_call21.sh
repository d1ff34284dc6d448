#!/bin/bash
mkdir -p gpurun_out
timeout 1700 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "emps" > gpurun_out/r02_gputests_21.log 2>&1
tail -n 4 gpurun_out/r02_gputests_21.log | cut -c1-300
timeout 600 python tools/gpu_probe.py x > gpurun_out/r02_microbench_21.log 2>&1; tail -2 gpurun_out/r02_microbench_21.log
timeout 600 python tools/emps_probe.py 2484 > gpurun_out/r02_emps_probe_21.log 2>&1; tail -4 gpurun_out/r02_emps_probe_21.log
PGAS_SWEEP_FUSED=1 timeout 600 python tools/emps_probe.py 2484 > gpurun_out/r02_emps_probe_fused_21.log 2>&1; tail -4 gpurun_out/r02_emps_probe_fused_21.log
timeout 900 python bench.py --config 5 --steps 2 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); r=d['roofline']; print('cfg5 ms_per_step',d['ms_per_step'],'value',d['value'],'state frac',r['frac'],'sweep_ms',r['sweep_ms'],'sweep_frac',r['sweep_frac']); print(d['roofline_kernels'])"
