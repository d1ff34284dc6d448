"""Developer aid: difflib similarity of repo .py files against the reference tree
(only meaningful in the build container where /root/reference exists)."""
import difflib, pathlib, sys
ref = [p for p in pathlib.Path("/root/reference").rglob("*.py")]
ref_txt = {p: p.read_text().splitlines() for p in ref}
root = pathlib.Path(__file__).resolve().parents[1]
worst = []
for p in root.rglob("*.py"):
    if "baseline/_ref" in str(p) or ".git/" in str(p):
        continue
    a = p.read_text().splitlines()
    for q, b in ref_txt.items():
        r = difflib.SequenceMatcher(None, a, b, autojunk=False).ratio()
        worst.append((r, str(p.relative_to(root)), str(q)))
worst.sort(reverse=True)
for r, p, q in worst[: int(sys.argv[1]) if len(sys.argv) > 1 else 10]:
    print(f"{r:.2f}  {p}  ~  {q}")
