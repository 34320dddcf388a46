"""Static opcode histogram per kernel of a cubin/.so: python tools/sass_stats.py lib.so [name-substring]"""
import collections, re, subprocess, sys
so = sys.argv[1]; pat = sys.argv[2] if len(sys.argv) > 2 else ""
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
fn = None; stats = collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s+Function : (\S+)", line)
    if m:
        fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()[:110]
        stats[fn] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and fn: stats[fn][m.group(1)] += 1
for fn, c in stats.items():
    if pat in fn:
        print(f"{fn}\n  total {sum(c.values())}: " + ", ".join(f"{o} {n}" for o, n in c.most_common(22)))
