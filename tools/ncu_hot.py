"""Top SASS instructions by stall samples from an .ncu-rep: python tools/ncu_hot.py rep [n]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = raw.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
def f(r, k):
    try: return float(r[k])
    except Exception: return 0.0
tot = sum(f(r, "# Samples") for r in rows) or 1
print("columns:", [k for k in rows[0].keys()][:12])
idx = sorted(range(len(rows)), key=lambda i: -f(rows[i], "# Samples"))[:n]
for i in sorted(idx):
    r = rows[i]
    print(f"{i:5d} {100*f(r,'# Samples')/tot:5.1f}%  exec {f(r,'Instructions Executed'):.3g}  {r['Source'][:100]}")
