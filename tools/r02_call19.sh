#!/bin/bash
# round-2 evidence run: launch list of the bench command, full-set captures of the dominant kernels (exported to text on the box:
# the .ncu-rep files are too large to travel), the sanitizer's answer on this pool
mkdir -p gpurun_out
O=gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-marginalised --no-cpu-baseline"
timeout 600 $B > $O/r02_bench_plain19.log 2> $O/r02_bench_plain19.err &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/r02_launches.csv $B > $O/r02_ncu_launches.log 2>&1
cap() {   # name, kernel regex, skip, count, command...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  timeout 600 "$@" > $O/r02_${name}_plain.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$rx" -s $skip -c $cnt -f -o /tmp/r02_$name "$@" > $O/r02_ncu_${name}.log 2>&1
  ncu -i /tmp/r02_$name.ncu-rep --page raw --csv > $O/r02_${name}_raw.csv 2>/dev/null
  python tools/ncu_top.py /tmp/r02_$name.ncu-rep 30 > $O/r02_${name}_top.txt 2>&1
}
cap state64 'csmc_state' 6 1 python tools/prof_sweep.py smo 4096 101 256 64 0
cap weights1 'csmc_weights1' 1 1 python tools/prof_sweep.py smo 4096 101 256 64 0
cap state8 'csmc_state' 10 1 python tools/prof_sweep.py smo 4096 201 256 8 0
cap wlat8 'csmc_weights_lat' 1 1 python tools/prof_sweep.py smo 4096 201 256 8 0
cap tail 'suffstats|chol_|mniw_draw' 14 40 python tools/tail_probe.py smo 256 2000 64 vehicle 1024 5000 16
compute-sanitizer --tool memcheck python -c "print(1)" > $O/r02_sanitizer.log 2>&1
ls -la $O | head -40; du -sh $O
