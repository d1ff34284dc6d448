#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02_gputests_2.log 2>&1
timeout 600 python tools/tail_probe.py smo 256 2000 64 emps 729 2484 1 emps 729 2484 8 vehicle 1024 5000 1 vehicle 1024 5000 8 > gpurun_out/r02_tail_probe_2.log 2>&1
PGAS_DRAW_BLOCKED=0 timeout 600 python tools/tail_probe.py smo 256 2000 64 > gpurun_out/r02_tail_probe_2b.log 2>&1
tail -n 4 gpurun_out/r02_gputests_2.log gpurun_out/r02_tail_probe_2.log gpurun_out/r02_tail_probe_2b.log
