#!/bin/bash
mkdir -p gpurun_out
P=bayesian_inference_with_explicit_and_implicit_prior_knowledge_b200
cp $P/libpgas_b200.so /tmp/default.so
out=gpurun_out/r02_variants_23.log; : > $out
for v in default rolled; do
  if [ $v = default ]; then cp /tmp/default.so $P/libpgas_b200.so; else cp $P/variants/libpgas_b200_$v.so $P/libpgas_b200.so; fi
  echo "== $v" >> $out
  timeout 600 python tools/state_probe.py 64 401 4 2>&1 | tail -n 1 | cut -c1-200 >> $out
  timeout 600 python tools/state_probe.py 16 201 5 2>&1 | tail -n 1 | cut -c1-200 >> $out
done
cp /tmp/default.so $P/libpgas_b200.so
cat $out
timeout 600 python tools/tail_probe.py smo 256 2000 64 vehicle 1024 5000 16 > gpurun_out/r02_tail_plain23.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'suffstats' -s 2 -c 2 -f -o /tmp/r02_suff python tools/tail_probe.py smo 256 2000 64 vehicle 1024 5000 16 > gpurun_out/r02_ncu_suff.log 2>&1
ncu -i /tmp/r02_suff.ncu-rep --page raw --csv > gpurun_out/r02_suffstats_kernel_raw.csv 2>/dev/null
python tools/ncu_top.py /tmp/r02_suff.ncu-rep 25 > gpurun_out/r02_suffstats_kernel_top.txt 2>&1
head -30 gpurun_out/r02_suffstats_kernel_top.txt | cut -c1-150
