#!/bin/bash
# likelihood_fcn as an expression program (fused kernel): parity cases first, then the whole GPU suite, smoke and the bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "pluginlik" 2>&1 | grep -E "^E  |passed|failed|FAILED|Error" | head -n 30 | cut -c1-300 | tee gpurun_out/r02_pluginlik_tests.txt
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -n 4 | cut -c1-300 | tee gpurun_out/r02_head_gputests.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2 | cut -c1-400
timeout 900 python bench.py --no-marginalised --no-cpu-baseline > gpurun_out/r02_head2_bench_cfg4.json 2> gpurun_out/r02_head2_bench_cfg4.err; echo "bench4 rc=$?"
python - <<'P'
import json
for line in open("gpurun_out/r02_head2_bench_cfg4.json"):
    if line.startswith("{"):
        d = json.loads(line); r = d["roofline"]
        print("ms %.2f value %.3e e2e %.3e frac %.3f sweep %.3f launches %d share8 %.2f" % (d["ms_per_step"], d["value"], d["e2e"]["value"], r["frac"], r["sweep_frac"], d["gpu_launches"], d["split_8gpu_share"]["ms_per_step"]))
P
