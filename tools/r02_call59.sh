#!/bin/bash
# final validation of round 2's HEAD: GPU suite, smoke, bench line
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
