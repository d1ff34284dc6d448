#!/bin/bash
# final suffstats kernel (warp-specialised SYRK): timings at both configuration shapes + one full-set capture each
mkdir -p gpurun_out
timeout 600 python tools/tail_probe.py smo 256 2000 64 vehicle 1024 5000 16 emps 729 2484 1 2>&1 | tail -n 4 | cut -c1-420 | tee gpurun_out/r02_final_tail_probe.txt
cap() {   # name, kernel regex, skip, count, command...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$rx" -s $skip -c $cnt -f -o /tmp/r02_$name "$@" > gpurun_out/r02_final_ncu_${name}.log 2>&1
  ncu -i /tmp/r02_$name.ncu-rep --page raw --csv > gpurun_out/r02_final_${name}_raw.csv 2>/dev/null
  python tools/ncu_top.py /tmp/r02_$name.ncu-rep 25 > gpurun_out/r02_final_${name}_top.txt 2>&1
  head -n 24 gpurun_out/r02_final_${name}_top.txt | cut -c1-160
}
cap suff_cfg4 'suffstats_kernel' 2 1 python tools/tail_probe.py smo 256 2000 64
cap suff_cfg5 'suffstats_kernel' 2 1 python tools/tail_probe.py vehicle 1024 5000 16
