#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_cfg4_a.log 2> gpurun_out/r02_bench_cfg4_a.err
timeout 900 python bench.py --config 5 --steps 2 --warmup 3 > gpurun_out/r02_bench_cfg5_a.log 2> gpurun_out/r02_bench_cfg5_a.err
timeout 600 python tools/tail_probe.py smo 256 2000 64 vehicle 1024 5000 8 > gpurun_out/r02_tail_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'suffstats|chol_|mniw_draw' -c 24 -o gpurun_out/r02_tail_a python tools/tail_probe.py smo 256 2000 64 vehicle 1024 5000 8 > gpurun_out/r02_ncu_tail.log 2>&1
tail -c 600 gpurun_out/r02_bench_cfg4_a.log gpurun_out/r02_bench_cfg5_a.log; tail -3 gpurun_out/r02_bench_cfg4_a.err gpurun_out/r02_bench_cfg5_a.err gpurun_out/r02_ncu_tail.log
