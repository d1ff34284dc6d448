#!/bin/bash
mkdir -p gpurun_out
timeout 1700 python -m pytest tests/test_gpu_marginal.py tests/test_drivers.py -q -m gpu --durations=12 > gpurun_out/r02_gputests_10.log 2>&1
tail -n 30 gpurun_out/r02_gputests_10.log
