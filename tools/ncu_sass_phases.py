"""Splits a kernel's executed warp-instructions into segments delimited by BAR.SYNC (and lists the
hottest opcodes), from `ncu -i X.ncu-rep --page source --csv`.  usage: ncu_sass_phases.py rep [particles]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; P = float(sys.argv[2]) if len(sys.argv) > 2 else 67108864.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = raw.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rd = csv.DictReader(io.StringIO("\n".join(lines[start:])))
seg, segs, ops, samples = 0, collections.OrderedDict(), collections.Counter(), collections.Counter()
tot = 0
for r in rd:
    try: n = float(r["Instructions Executed"])
    except Exception: continue
    src = r["Source"].strip()
    op = src.split()[0] if not src.startswith("@") else src.split()[1]
    op = op.split(".")[0]
    segs[seg] = segs.get(seg, 0) + n
    ops[op] += n; tot += n
    try: samples[seg] += float(r["# Samples"])
    except Exception: pass
    if "BAR.SYNC" in src or src.startswith("BAR"): seg += 1
print(f"total warp-instr {tot:.4g} = {tot*32/P:.0f} thread-instr/particle")
ts = sum(samples.values()) or 1
for s, n in segs.items(): print(f"  segment {s}: {n*32/P:8.1f} instr/particle  {100*samples[s]/ts:5.1f}% of stall samples")
print("  top opcodes:", ", ".join(f"{o} {n*32/P:.0f}" for o, n in ops.most_common(14)))
