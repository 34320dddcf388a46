"""A/B harness: builds kernel variants (compile-time -D switches) into gpurun_out-free temp libs
here, or — with --run — times each prebuilt variant on the GPU box in its own process.
  python tools/ab.py build name1:DEF=1,DEF2=0 name2:...      (CPU box)
  python tools/ab.py run name1 name2 ...                     (GPU box)"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VAR = os.path.join(ROOT, "mpm_b200", "variants")

if sys.argv[1] == "build":
    from mpm_b200 import build as b
    os.makedirs(VAR, exist_ok=True)
    for spec in sys.argv[2:]:
        name, _, defs = spec.partition(":")
        defines = tuple(d for d in defs.split(",") if d)
        out = os.path.join(VAR, f"libmpm_{name}.so")
        b.build(force=True, defines=defines, out=out)
        print("built", out, defines)
else:
    extra = os.environ.get("AB_BENCH_ARGS", "--steps 16 --warmup 8 --no-e2e --no-cpu --no-extras").split()
    for name in sys.argv[2:]:
        env = dict(os.environ, MPM_B200_LIB=os.path.join(VAR, f"libmpm_{name}.so"))
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + extra, env=env, capture_output=True, text=True)
        try:
            d = json.loads(r.stdout.strip().splitlines()[-1])
            st = d["stage_ms"]
            print(f"{name:24s} step {d['ms_per_step']:.3f} ms | p2g {st['p2g']:.3f} g2p {st['g2p']:.3f} sort {st['sort']:.3f} grid {st['grid']:.3f} reset {st['reset']:.3f} | frac {d['substep_roofline']['frac']:.3f}", flush=True)
        except Exception as e:
            print(name, "FAILED", e, r.stderr[-800:], flush=True)
