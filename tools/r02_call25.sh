#!/bin/bash
mkdir -p gpurun_out
P=bayesian_inference_with_explicit_and_implicit_prior_knowledge_b200
cp $P/libpgas_b200.so /tmp/default.so
run() { env $2 timeout 600 python bench.py --steps 3 --warmup 3 --no-marginalised --no-cpu-baseline --no-strong 2>/dev/null | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); r=d['roofline']; print('[$1 $2] ms_per_step %.2f state frac %.4f sweep_ms %.2f' % (d['ms_per_step'], r['frac'], r['sweep_ms']))"; }
for e in "A=1" "PGAS_SPLIT_STATE_ROWS=8" "PGAS_SPLIT_STATE_ROWS=32" "PGAS_SPLIT_STATE_ROWS=64" "PGAS_SPLIT_ROWS=32" "PGAS_SPLIT_ONE_GROUP=1"; do run default "$e"; done
for v in nt128 g4 g1; do cp $P/variants/libpgas_b200_$v.so $P/libpgas_b200.so; run $v "A=1"; done
cp /tmp/default.so $P/libpgas_b200.so
