#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_pins.py tests/test_drivers.py -x -q -m gpu 2>&1 | tail -n 4 | cut -c1-300
timeout 600 python tools/emps_probe.py 400 1 2>&1 | head -2 | cut -c1-200
PGAS_STATE_LANES=0 timeout 600 python tools/emps_probe.py 400 1 2>&1 | head -1 | cut -c1-200
PGAS_SPLIT_TIMELINE=1 timeout 600 python tools/emps_probe.py 400 1 2>&1 | grep -A8 "^chunk" | head -9
