#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_marginal.py -q -m gpu 2>&1 | grep -E "^E  |passed|failed|FAILED" | head -n 30 | cut -c1-250
for w in smo vehicle; do timeout 300 python tools/marg_probe.py $w 4 1 2>&1 | tail -n 2 | cut -c1-260; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:marg_sweep -c 1 -f -o gpurun_out/r02_marg_b python tools/prof_marg.py smo 101 200 41 > gpurun_out/ncu_marg.log 2>&1
python tools/ncu_lines.py gpurun_out/r02_marg_b.ncu-rep 40 samp 2>&1 | cut -c1-200 | head -n 12
