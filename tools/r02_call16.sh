#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/prof_sweep.py smo 4096 201 256 8 0 > gpurun_out/r02_prof_plain16.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:csmc_weights_lat -s 1 -c 1 -f -o gpurun_out/r02_wlat_a python tools/prof_sweep.py smo 4096 201 256 8 0 > gpurun_out/r02_ncu_wlat.log 2>&1
tail -n 3 gpurun_out/r02_prof_plain16.log gpurun_out/r02_ncu_wlat.log
