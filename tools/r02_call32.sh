#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02_bench_2gpu.log 2> gpurun_out/r02_bench_2gpu.err
python - <<'PY'
import json
for line in open('gpurun_out/r02_bench_2gpu.log'):
    if line.startswith('{'):
        d=json.loads(line); print({k:d[k] for k in ('value','n_gpus','ms_per_step','scaling')}); print(d.get('strong'))
PY
tail -n 3 gpurun_out/r02_bench_2gpu.err | cut -c1-200
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 2>/dev/null | tail -c 400
