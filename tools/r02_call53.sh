#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_marginal.py -x -q -m gpu 2>&1 | tail -n 5 | cut -c1-300
for w in smo vehicle emps; do timeout 300 python tools/marg_probe.py $w 4 1,7 2>&1 | tail -n 3 | cut -c1-260; done
PGAS_MARG_NARROW=1 timeout 300 python tools/marg_probe.py smo 4 1 2>&1 | tail -n 1 | cut -c1-260
