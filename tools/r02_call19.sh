#!/bin/bash
# round-2 evidence run: launch list of the bench command, full-set captures of the dominant kernels, compute-sanitizer on smoke()
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-marginalised --no-cpu-baseline"
timeout 600 $B > gpurun_out/r02_bench_plain19.log 2> gpurun_out/r02_bench_plain19.err &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_launches.csv $B > gpurun_out/r02_ncu_launches.log 2>&1
timeout 600 python tools/prof_sweep.py smo 4096 101 256 64 0 > gpurun_out/r02_prof64_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'csmc_state|csmc_weights1' -s 6 -c 3 -f -o gpurun_out/r02_sweep64 python tools/prof_sweep.py smo 4096 101 256 64 0 > gpurun_out/r02_ncu_sweep64.log 2>&1
timeout 600 python tools/prof_sweep.py smo 4096 201 256 8 0 > gpurun_out/r02_prof8_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'csmc_state|csmc_weights_lat' -s 10 -c 4 -f -o gpurun_out/r02_sweep8 python tools/prof_sweep.py smo 4096 201 256 8 0 > gpurun_out/r02_ncu_sweep8.log 2>&1
timeout 600 python tools/tail_probe.py smo 256 2000 64 vehicle 1024 5000 16 > gpurun_out/r02_tail_plain19.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'suffstats|chol_|mniw_draw' -s 14 -c 40 -f -o gpurun_out/r02_tail python tools/tail_probe.py smo 256 2000 64 vehicle 1024 5000 16 > gpurun_out/r02_ncu_tail.log 2>&1
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_sanitizer_memcheck.log 2>&1
timeout 1200 compute-sanitizer --tool racecheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_sanitizer_racecheck.log 2>&1
tail -n 4 gpurun_out/r02_bench_plain19.err gpurun_out/r02_ncu_launches.log gpurun_out/r02_ncu_sweep64.log gpurun_out/r02_ncu_sweep8.log gpurun_out/r02_ncu_tail.log gpurun_out/r02_tail_plain19.log
tail -n 6 gpurun_out/r02_sanitizer_memcheck.log gpurun_out/r02_sanitizer_racecheck.log
