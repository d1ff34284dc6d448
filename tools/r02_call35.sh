#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:csmc_state -s 6 -c 1 -f -o gpurun_out/r02_state_b python tools/prof_sweep.py smo 4096 101 256 64 0 > gpurun_out/r02_ncu_state_b.log 2>&1
tail -n 2 gpurun_out/r02_ncu_state_b.log
ls -la gpurun_out/r02_state_b.ncu-rep
