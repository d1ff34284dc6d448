#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "suffstats or blocked or draw or run_chains or posterior_predictive" 2>&1 | tail -n 3 | cut -c1-250
timeout 600 python tools/tail_probe.py smo 256 2000 64 vehicle 1024 5000 16 emps 729 2484 1 smo 256 2000 8 2>&1 | tail -n 4 | cut -c1-420
