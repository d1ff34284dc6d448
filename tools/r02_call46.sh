#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python bench.py ) > gpurun_out/r02_bench_default_46.log 2> gpurun_out/r02_bench_default_46.err; tail -n 4 gpurun_out/r02_bench_default_46.err
( time timeout 900 python bench.py --config 5 --steps 2 --warmup 3 --no-marginalised --no-cpu-baseline --no-strong ) > gpurun_out/r02_bench_cfg5_46.log 2> gpurun_out/r02_bench_cfg5_46.err; tail -n 4 gpurun_out/r02_bench_cfg5_46.err
