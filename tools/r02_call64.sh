#!/bin/bash
# 2-GPU run of the final code (torchrun, NCCL): weak scaling + the "strong" record
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02_final_bench_2gpu.json 2> gpurun_out/r02_final_bench_2gpu.err; echo "rc=$?"
python - <<'P'
import json
for line in open("gpurun_out/r02_final_bench_2gpu.json"):
    if line.startswith("{"):
        d = json.loads(line)
        print("n_gpus %d ms %.2f value %.3e e2e %.3e" % (d["n_gpus"], d["ms_per_step"], d["value"], d["e2e"]["value"]))
        print("strong", d.get("strong"))
P
tail -n 3 gpurun_out/r02_final_bench_2gpu.err | cut -c1-300
