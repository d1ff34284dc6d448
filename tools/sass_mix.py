"""Static SASS instruction mix of one kernel: python tools/sass_mix.py <object> <mangled-substring> [lo hi]
Lists backward branches (loops) and the opcode histogram of the whole function or of the byte range [lo, hi]."""
import collections
import re
import subprocess
import sys

obj, pat = sys.argv[1], sys.argv[2]
lo = int(sys.argv[3], 16) if len(sys.argv) > 3 else 0
hi = int(sys.argv[4], 16) if len(sys.argv) > 4 else 1 << 60
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
cur, rows = None, []
for line in txt.splitlines():
    m = re.match(r"\s+Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    if cur is None or pat not in cur:
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m:
        rows.append((int(m.group(1), 16), m.group(2).strip()))
print("instructions:", len(rows))
for a, s in rows:
    m = re.search(r"\bBRA\S*\s+(?:\S+,\s*)?(0x[0-9a-f]+)", s)
    if m and int(m.group(1), 16) <= a:
        print("loop  %05x -> %05x  (%d instr)  %s" % (a, int(m.group(1), 16), (a - int(m.group(1), 16)) // 16 + 1, s))
h = collections.Counter()
for a, s in rows:
    if lo <= a <= hi:
        t = s.split()
        op = t[1] if t[0].startswith("@") else t[0]
        h[op.split(".")[0]] += 1
tot = sum(h.values())
print("range %x..%x: %d instr" % (lo, min(hi, rows[-1][0]), tot))
for k, v in h.most_common():
    print("  %5d  %s" % (v, k))
