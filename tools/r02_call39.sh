#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'suffstats' -s 2 -c 2 -f -o gpurun_out/r02_suff_b python tools/tail_probe.py smo 256 2000 64 vehicle 1024 5000 16 > gpurun_out/r02_ncu_suff_b.log 2>&1
tail -n 2 gpurun_out/r02_ncu_suff_b.log | cut -c1-200
