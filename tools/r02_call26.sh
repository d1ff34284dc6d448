#!/bin/bash
mkdir -p gpurun_out
( time timeout 1700 python -m pytest tests -x -q -m gpu ) > gpurun_out/r02_gputests_26.log 2>&1
tail -n 8 gpurun_out/r02_gputests_26.log | cut -c1-200
( time timeout 600 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r02_smoke_26.log 2>&1; tail -n 5 gpurun_out/r02_smoke_26.log | cut -c1-300
( time timeout 900 python bench.py ) > gpurun_out/r02_bench_default_26.log 2> gpurun_out/r02_bench_default_26.err; tail -n 4 gpurun_out/r02_bench_default_26.err; tail -c 600 gpurun_out/r02_bench_default_26.log
( time timeout 900 python bench.py --impl reference ) > gpurun_out/r02_bench_reference_26.log 2> gpurun_out/r02_bench_reference_26.err; tail -n 4 gpurun_out/r02_bench_reference_26.err; tail -c 800 gpurun_out/r02_bench_reference_26.log
