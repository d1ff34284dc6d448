#!/bin/bash
# final-state evidence of round 2: GPU suite, smoke, both bench lines, launch list, marginalised / EMPS probes
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -n 4 | cut -c1-300 | tee gpurun_out/r02_final_gputests.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2 | cut -c1-400
timeout 900 python bench.py > gpurun_out/r02_final_bench_cfg4.json 2> gpurun_out/r02_final_bench_cfg4.err; echo "bench4 rc=$?"
timeout 900 python bench.py --config 5 --steps 2 --warmup 3 --no-marginalised > gpurun_out/r02_final_bench_cfg5.json 2> gpurun_out/r02_final_bench_cfg5.err; echo "bench5 rc=$?"
python - <<'P'
import json
for f in ("gpurun_out/r02_final_bench_cfg4.json", "gpurun_out/r02_final_bench_cfg5.json"):
    for line in open(f):
        if line.startswith("{"):
            d = json.loads(line); r = d["roofline"]
            print(f, "ms %.2f value %.3e e2e %.3e frac %.3f" % (d["ms_per_step"], d["value"], d["e2e"]["value"], r["frac"]))
            print({k: r[k] for k in r if k != "traffic_capture"})
            for k in ("roofline_kernels", "split_8gpu_share", "marginalised", "cpu_baseline", "clocks"):
                print(k, d.get(k))
P
for w in smo vehicle emps; do timeout 300 python tools/marg_probe.py $w 4 1,7 2>&1 | tail -n 3 | cut -c1-260; done | tee gpurun_out/r02_final_marg_probe.txt
timeout 300 python tools/emps_probe.py 2>&1 | tail -n 4 | cut -c1-300 | tee gpurun_out/r02_final_emps_probe.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_final_launches.csv python bench.py --steps 1 --warmup 1 --no-marginalised --no-cpu-baseline > gpurun_out/r02_final_ncu.log 2>&1; echo "ncu rc=$?"
cap() {   # name, kernel regex, skip, count, command...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$rx" -s $skip -c $cnt -f -o /tmp/r02_$name "$@" > gpurun_out/r02_final_ncu_${name}.log 2>&1
  ncu -i /tmp/r02_$name.ncu-rep --page raw --csv > gpurun_out/r02_final_${name}_raw.csv 2>/dev/null
  python tools/ncu_top.py /tmp/r02_$name.ncu-rep 30 > gpurun_out/r02_final_${name}_top.txt 2>&1
  head -n 24 gpurun_out/r02_final_${name}_top.txt | cut -c1-160
}
cap state64 'csmc_state' 6 1 python tools/prof_sweep.py smo 4096 101 256 64 0
cap marg 'marg_sweep' 0 1 python tools/prof_marg.py smo 101 200 41
