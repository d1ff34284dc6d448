#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/prof_sweep.py smo 4096 101 256 64 2 > gpurun_out/r02_prof_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:csmc_state -s 4 -c 2 -f -o gpurun_out/r02_state_a python tools/prof_sweep.py smo 4096 101 256 64 2 > gpurun_out/r02_ncu_state.log 2>&1
tail -n 3 gpurun_out/r02_prof_plain.log gpurun_out/r02_ncu_state.log
