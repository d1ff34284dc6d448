#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_pins.py -x -q -m gpu > gpurun_out/r02_gputests_7.log 2>&1
timeout 600 python tools/state_probe.py 64 401 4 > gpurun_out/r02_state_probe_7.log 2>&1
timeout 600 python tools/state_probe.py 16 201 5 >> gpurun_out/r02_state_probe_7.log 2>&1
timeout 900 python bench.py --steps 3 --warmup 3 --no-marginalised --no-cpu-baseline > gpurun_out/r02_bench_cfg4_7.log 2> gpurun_out/r02_bench_cfg4_7.err
tail -n 3 gpurun_out/r02_gputests_7.log gpurun_out/r02_state_probe_7.log
python - <<'PY'
import json
for line in open("gpurun_out/r02_bench_cfg4_7.log"):
    if line.startswith("{"):
        d=json.loads(line); r=d["roofline"]
        print("ms_per_step",d["ms_per_step"],"value",d["value"],"state frac",r["frac"],"sweep_ms",r["sweep_ms"],"sweep_frac",r["sweep_frac"], r["note"][-120:])
PY
