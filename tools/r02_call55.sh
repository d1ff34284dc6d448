#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_marginal.py -q -m gpu 2>&1 | grep -E "^E  |passed|failed|FAILED" | head -n 30 | cut -c1-250
for w in smo vehicle emps; do timeout 300 python tools/marg_probe.py $w 4 1,7 2>&1 | tail -n 3 | cut -c1-260; done
