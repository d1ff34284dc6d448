#!/bin/bash
mkdir -p gpurun_out
PGAS_SPLIT_TIMELINE=1 timeout 600 python tools/prof_sweep.py smo 4096 2000 256 8 0 > gpurun_out/r02_timeline_17.log 2>&1
tail -40 gpurun_out/r02_timeline_17.log
