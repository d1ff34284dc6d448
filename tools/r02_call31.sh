#!/bin/bash
mkdir -p gpurun_out
PGAS_SPLIT_TIMELINE=1 timeout 600 python tools/emps_probe.py 700 > gpurun_out/r02_emps_timeline.log 2>&1
grep -A14 "chunk:" gpurun_out/r02_emps_timeline.log | tail -16
PGAS_SPLIT_SERIAL=1 timeout 600 python tools/emps_probe.py 2484 2>&1 | head -1
