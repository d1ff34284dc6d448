#!/bin/bash
# A/B of compile-time variants of the state kernel (built into variants/ by the developer): state kernel alone + bench iteration
mkdir -p gpurun_out
P=bayesian_inference_with_explicit_and_implicit_prior_knowledge_b200
cp $P/libpgas_b200.so /tmp/default.so
out=gpurun_out/r02_variants_9.log; : > $out
for v in default pp1 u1 u4 pp1u4; do
  if [ $v = default ]; then cp /tmp/default.so $P/libpgas_b200.so; else cp $P/variants/libpgas_b200_$v.so $P/libpgas_b200.so; fi
  echo "== $v" >> $out
  timeout 600 python tools/state_probe.py 64 401 4 2>&1 | tail -n 1 | cut -c1-330 >> $out
  timeout 600 python tools/state_probe.py 16 201 5 2>&1 | tail -n 1 | cut -c1-330 >> $out
  timeout 900 python bench.py --steps 2 --warmup 3 --no-marginalised --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); r=d['roofline']; print('bench ms_per_step',d['ms_per_step'],'state frac',r['frac'],'sweep_ms',r['sweep_ms'])" >> $out
done
cp /tmp/default.so $P/libpgas_b200.so
cat $out
