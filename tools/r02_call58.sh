#!/bin/bash
# model plug-in of the marginalised path (expression programs): parity cases + regression check of the table-driven kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_marginal.py -q -m gpu 2>&1 | grep -E "^E  |passed|failed|FAILED|Error" | head -n 40 | cut -c1-300 | tee gpurun_out/r02_plugin_tests.txt
for w in smo vehicle emps; do timeout 300 python tools/marg_probe.py $w 4 1 2>&1 | tail -n 2 | cut -c1-260; done | tee gpurun_out/r02_plugin_marg_probe.txt
