"""Executed instructions per particle by opcode (and by BAR-delimited segment) for one kernel of an
.ncu-rep: python tools/ncu_opcodes.py rep kernel_regex [particles]"""
import collections, csv, io, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
P = float(sys.argv[3]) if len(sys.argv) > 3 else 67108864.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
lines = raw.splitlines()
starts = [i for i, l in enumerate(lines) if l.startswith('"Address"')]
s = starts[0]
e = starts[1] if len(starts) > 1 else len(lines)
rd = csv.DictReader(io.StringIO("\n".join(lines[s:e])))
seg, segops, samples = 0, collections.defaultdict(collections.Counter), collections.Counter()
for r in rd:
    try:
        n = float(r["Instructions Executed"])
    except Exception:
        continue
    src = r["Source"].strip()
    toks = src.split()
    op = toks[1] if src.startswith("@") else toks[0]
    op = ".".join(op.split(".")[:2]) if op.startswith(("LDS", "STS", "LDG", "STG", "RED", "ATOM")) else op.split(".")[0]
    segops[seg][op] += n
    try:
        samples[seg] += float(r["# Samples"])
    except Exception:
        pass
    if src.startswith("BAR") or " BAR.SYNC" in src:
        seg += 1
ts = sum(samples.values()) or 1
tot = sum(sum(c.values()) for c in segops.values())
print(f"total {tot * 32 / P:.0f} warp-instr x 32 / particle")
for sg, c in segops.items():
    t = sum(c.values())
    print(f"seg {sg}: {t * 32 / P:.1f}/particle, {100 * samples[sg] / ts:.0f}% of samples:", ", ".join(f"{o} {n * 32 / P:.1f}" for o, n in c.most_common(24)))
