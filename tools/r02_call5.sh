#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/state_probe.py 64 401 4 > gpurun_out/r02_state_probe_a.log 2>&1
timeout 900 python tools/state_probe.py 16 201 5 >> gpurun_out/r02_state_probe_a.log 2>&1
timeout 600 python tools/tail_probe.py smo 256 2000 64 vehicle 1024 5000 16 > gpurun_out/r02_tail_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'suffstats|chol_|mniw_draw' -s 14 -c 40 -o gpurun_out/r02_tail_b python tools/tail_probe.py smo 256 2000 64 vehicle 1024 5000 16 > gpurun_out/r02_ncu_tail_b.log 2>&1
tail -n 3 gpurun_out/r02_state_probe_a.log gpurun_out/r02_ncu_tail_b.log
