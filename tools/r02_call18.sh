#!/bin/bash
mkdir -p gpurun_out
timeout 1700 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_pins.py -q -m gpu > gpurun_out/r02_gputests_18.log 2>&1
tail -n 12 gpurun_out/r02_gputests_18.log | cut -c1-300
