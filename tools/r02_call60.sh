#!/bin/bash
# A/B of compile-time variants at configs[3] (64 chains): state kernel with 128-thread CTAs, one-CTA resampling kernel at 32 registers
# (so that three 128-thread state CTAs fit next to a resampling CTA instead of one 256-thread CTA)
mkdir -p gpurun_out
P=bayesian_inference_with_explicit_and_implicit_prior_knowledge_b200/build_variants
for v in "" $P/libB.so $P/libC.so $P/libD.so ""; do
PGAS_LIB_PATH=$v timeout 600 python bench.py --steps 3 --warmup 3 --no-marginalised --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); r=d['roofline']; print('[${v:-default}] ms_per_step %.2f state frac %.4f sweep_ms %.2f share8 %.2f' % (d['ms_per_step'], r['frac'], r['sweep_ms'], d['split_8gpu_share']['ms_per_step']))"
done | tee gpurun_out/r02_variants_nt128.txt
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -n 4 | cut -c1-300 | tee gpurun_out/r02_head_gputests.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2 | cut -c1-400
timeout 900 python bench.py > gpurun_out/r02_head_bench_cfg4.json 2> gpurun_out/r02_head_bench_cfg4.err; echo "bench4 rc=$?"
python - <<'P'
import json
for line in open("gpurun_out/r02_head_bench_cfg4.json"):
    if line.startswith("{"):
        d = json.loads(line); r = d["roofline"]
        print("ms %.2f value %.3e e2e %.3e frac %.3f sweep %.3f launches %d" % (d["ms_per_step"], d["value"], d["e2e"]["value"], r["frac"], r["sweep_frac"], d["gpu_launches"]))
        print(d["marginalised"]["chains_1"], d["split_8gpu_share"]["ms_per_step"])
P
