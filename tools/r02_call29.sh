#!/bin/bash
mkdir -p gpurun_out
P=bayesian_inference_with_explicit_and_implicit_prior_knowledge_b200
cp $P/libpgas_b200.so /tmp/default.so
cp $P/variants/libpgas_b200_nt512p4.so $P/libpgas_b200.so
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "split_and_fused and 3]" 2>&1 | tail -3
for e in "PGAS_WEIGHTS_KERNEL=3" "PGAS_WEIGHTS_KERNEL=3 PGAS_SPLIT_STATE_ROWS=8"; do
env $e timeout 900 python bench.py --config 5 --steps 2 --warmup 2 --no-cpu-baseline --no-strong 2>/dev/null | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); r=d['roofline']; print('[nt512p4 $e] cfg5 ms_per_step %.1f value %.3e state frac %.3f sweep_ms %.1f sweep_frac %.3f' % (d['ms_per_step'],d['value'],r['frac'],r['sweep_ms'],r['sweep_frac']))"
done
cp /tmp/default.so $P/libpgas_b200.so
