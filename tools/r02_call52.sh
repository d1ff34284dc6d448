#!/bin/bash
# marginalised kernel: where do the 18 k warp instructions of a particle-step go? (per-line counts from one full capture)
timeout 300 python tools/prof_marg.py smo 101 200 41 2>&1 | tail -n 2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:marg_sweep -c 1 -f -o gpurun_out/r02_marg_a python tools/prof_marg.py smo 101 200 41 > gpurun_out/ncu_marg.log 2>&1
python tools/ncu_lines.py gpurun_out/r02_marg_a.ncu-rep 70 inst > gpurun_out/r02_marg_lines_inst.txt 2>&1
python tools/ncu_lines.py gpurun_out/r02_marg_a.ncu-rep 50 samp > gpurun_out/r02_marg_lines_samp.txt 2>&1
head -n 75 gpurun_out/r02_marg_lines_inst.txt | cut -c1-200
timeout 300 python tools/marg_probe.py smo 4 1,7 2>&1 | tail -n 4 | cut -c1-300
