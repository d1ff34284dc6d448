"""profiles/r02_sass_excerpts.md: per hot kernel the static opcode mix and a SASS excerpt around its densest FP64 / tensor / async region
(cuobjdump -sass of the in-tree objects; no GPU needed).  usage: python tools/sass_excerpts.py > profiles/r02_sass_excerpts.md"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B = os.path.join(ROOT, "bayesian_inference_with_explicit_and_implicit_prior_knowledge_b200", "build")
KERNELS = [("sweep.o", "csmc_state_kernelILi2ELi1ELb0ELi256ELi2", r"DFMA", "state kernel <n_x=2, n_y=1, Philox, 256 threads, 2 particles per thread>: the row walk"),
           ("weights_lat.o", "csmc_weights_lat_kernelILi256ELi2ELi8", r"SYNCS|STAS|ST\.ASYNC|UBLKCP", "latency form of the resampling kernel, cluster of 8: mbarrier waits and st.async hand-offs"),
           ("weights.o", "csmc_weights1_kernelILi512ELi4ELi2", r"BAR\.SYNC|DFMA", "one-CTA form of the resampling kernel"),
           ("suffstats.o", "suffstats_kernelILi2", r"DMMA|UBLKCP|SYNCS", "sufficient statistics: SYRK over time on the FP64 tensor pipe"),
           ("mniw_draw.o", "chol_update", r"DMMA", "blocked Cholesky, trailing update on the FP64 tensor pipe"),
           ("mniw_draw.o", "mniw_draw_kernel", r"DFMA", "posterior draw"),
           ("marginal.o", "marg_sweep_kernelILi1ELi2ELb1ELb0", r"DFMA", "marginalised conditional sweep (Algorithm3), 255-register instantiation of the wide geometry (the shipped single-mass oscillator runs this one)")]
print("# r02 — SASS evidence (cuobjdump -sass of the in-tree objects built for sm_100a)\n")
print("`tcgen05.mma` has no f64 kind: the Blackwell FP64 tensor instruction is `DMMA.8x8x4` (PTX `mma.sync.m8n8k4.f64`); it shares the FP64 pipe with `DFMA`")
print("(profiles/r02_microbench.md), so kernels whose contraction is n_x = 2 columns wide use DFMA and the dense ones (statistics, trailing updates) DMMA.\n")
for obj, pat, key, title in KERNELS:
    txt = subprocess.run(["cuobjdump", "-sass", os.path.join(B, obj)], capture_output=True, text=True).stdout
    cur, rows, name = None, [], None
    for line in txt.splitlines():
        m = re.match(r"\s+Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        if cur is None or pat not in cur:
            continue
        name = name or cur
        if cur != name:
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            rows.append((int(m.group(1), 16), m.group(2).strip()))
    if not rows:
        print(f"## {title}\n\n(kernel `{pat}` not found in {obj})\n")
        continue
    mix = collections.Counter((s.split()[1] if s.split()[0].startswith("@") else s.split()[0]).split(".")[0] for _, s in rows)
    print(f"## {title}\n\n`{name}` in `{obj}`: {len(rows)} instructions; " + ", ".join(f"{k} {v}" for k, v in mix.most_common(10)) + "\n")
    hits = [i for i, (_, s) in enumerate(rows) if re.search(key, s)]
    if hits:
        # densest window of 28 instructions
        best = max(range(0, max(1, len(rows) - 28)), key=lambda i: sum(1 for h in hits if i <= h < i + 28))
        print("```")
        for a, s in rows[best:best + 28]:
            print(f"/*{a:05x}*/  {s}")
        print("```\n")
    else:
        print(f"(no `{key}` instruction)\n")
