#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02_gputests_4.log 2>&1
timeout 600 python tools/tail_probe.py smo 256 2000 64 emps 729 2484 1 vehicle 1024 5000 1 vehicle 1024 5000 16 > gpurun_out/r02_tail_probe_4.log 2>&1
tail -n 5 gpurun_out/r02_gputests_4.log gpurun_out/r02_tail_probe_4.log
