#!/bin/bash
# round-2 GPU call 1: where does round 1's code stand on the shapes it never exercised?
mkdir -p gpurun_out
python tools/gpu_probe.py x > gpurun_out/r02_probe.log 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "suffstats_and_draw" > gpurun_out/r02_parity_tail.log 2>&1
timeout 600 python tools/tail_probe.py smo 256 2000 64 emps 729 2484 1 emps 729 2484 8 vehicle 1024 5000 1 vehicle 1024 5000 8 > gpurun_out/r02_tail_probe.log 2>&1
timeout 600 python tools/cfg5_probe.py 201 8 > gpurun_out/r02_cfg5_8.log 2>&1
timeout 600 python tools/cfg5_probe.py 201 16 > gpurun_out/r02_cfg5_16.log 2>&1
timeout 600 python bench.py --steps 3 --warmup 3 --chains 8 --no-marginalised --no-cpu-baseline > gpurun_out/r02_bench_8chains.log 2>&1
tail -3 gpurun_out/r02_*.log
