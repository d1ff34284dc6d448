#!/bin/bash
# round-2 call 6 (re-entry): state of the committed code — full GPU suite, both bench configs, tail probe
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02_gputests_6.log 2>&1
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_cfg4_6.log 2> gpurun_out/r02_bench_cfg4_6.err
timeout 900 python bench.py --config 5 --steps 2 --warmup 3 > gpurun_out/r02_bench_cfg5_6.log 2> gpurun_out/r02_bench_cfg5_6.err
timeout 600 python tools/tail_probe.py smo 256 2000 64 emps 729 2484 1 vehicle 1024 5000 1 vehicle 1024 5000 16 > gpurun_out/r02_tail_probe_6.log 2>&1
timeout 600 python bench.py --steps 3 --warmup 3 --chains 8 --no-marginalised --no-cpu-baseline > gpurun_out/r02_bench_8chains_6.log 2>&1
tail -n 5 gpurun_out/r02_gputests_6.log gpurun_out/r02_tail_probe_6.log
tail -c 1500 gpurun_out/r02_bench_cfg4_6.log gpurun_out/r02_bench_cfg5_6.log gpurun_out/r02_bench_8chains_6.log
tail -n 3 gpurun_out/*_6.err
