"""Aggregate an `ncu --page source --print-source cuda,sass --csv` export per source line (developer aid).
usage: ncu_lines.py report.ncu-rep [top] [sort: inst|samp]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
key = sys.argv[3] if len(sys.argv) > 3 else "inst"
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
agg, cur, hdr = {}, None, None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        ii, si = r.index("Instructions Executed"), r.index("# Samples")
        continue
    if hdr is None or r[0] in ("-", ""):
        continue
    try:
        ln = int(r[0])
    except ValueError:
        continue
    a = agg.setdefault((cur, ln, r[1][:110]), [0, 0])
    a[0] += int(r[ii]) if r[ii].isdigit() else 0
    a[1] += int(r[si]) if r[si].isdigit() else 0
ti, ts = sum(v[0] for v in agg.values()), sum(v[1] for v in agg.values())
byf = collections.Counter()
for k, v in agg.items():
    byf[k[0]] += v[0]
print("total inst", ti, "samples", ts, {k: round(100 * v / ti, 1) for k, v in byf.most_common(8)})
ix = 0 if key == "inst" else 1
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][ix])[:top]:
    print(f"{100 * v[0] / ti:5.1f}% inst {100 * v[1] / ts:5.1f}% samp  {k[0]}:{k[1]}  {k[2]}")
