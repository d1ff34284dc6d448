#!/bin/bash
timeout 600 ncu --set full --clock-control none --import-source on -k regex:marg_sweep -c 1 -f -o gpurun_out/r02_marg_c python tools/prof_marg.py smo 101 200 41 > gpurun_out/ncu_marg.log 2>&1
python tools/ncu_lines.py gpurun_out/r02_marg_c.ncu-rep 10 samp 2>&1 | cut -c1-200 | head -n 4
timeout 600 ncu --set full --clock-control none --import-source on -k regex:marg_sweep -c 1 -f -o gpurun_out/r02_marg_veh python tools/prof_marg.py vehicle 101 200 20 > gpurun_out/ncu_marg2.log 2>&1
python tools/ncu_lines.py gpurun_out/r02_marg_veh.ncu-rep 10 samp 2>&1 | cut -c1-200 | head -n 4
