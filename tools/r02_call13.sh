#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/ticks_bench.py 2 8 > gpurun_out/r02_ticks_13.log 2>&1
PGAS_SPLIT_SERIAL=1 timeout 600 python tools/ticks_bench.py 2 8 >> gpurun_out/r02_ticks_13.log 2>&1
cat gpurun_out/r02_ticks_13.log
