#!/usr/bin/env python
"""Benchmark of the MLS-MPM substep (BASELINE.json metric: particle-steps/s + HBM roofline).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N=1 workload = BASELINE.json configs[3]: synthetic dense block, N=256, 2^26 particles,
fixed-corotated.  For N>1 (launched with torch.distributed.run, one
rank per GPU) every rank keeps 2^26 particles and ~2^24 grid nodes (weak scaling): the cubic
domain grows to N = 256 * gpus^(1/3) and is cut into x-slabs balanced by particle count.
A "step" is one substep (grid reset -> P2G -> [halo exchange] -> grid update -> G2P), the re-bin is
included at its cadence (--sort-every).  Inputs are generated on the device and are far larger
than L2 (6.7 GB particles + 268 MB grid per GPU), so no L2 flush is needed between steps.

Outside the timed region the run checks itself (exit code 3 on failure): particle count, grid mass
after P2G, finiteness ("invariants"), and at N > 1 a small scene on the same N slabs against the CPU
checker ("parity_nranks").  Further keys of the JSON line: the other BASELINE.json configurations
("configs": snow = configs[4], strong scaling of configs[3]) and stressed variants of the workload
("stressed": shear flow + perturbed F, snow that yields, bit-exact svd3 mode).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

P_PER_GPU = 1 << 26
BYTES_PER_PARTICLE = 252   # P2G read 100 + G2P read 52 + G2P write 100
BYTES_PER_NODE = 80        # zero 16 + P2G write-back 16 + grid update 16+16 + G2P read 16
# per-kernel algorithmic bytes (DESIGN.md "Kernels"): (bytes per particle, bytes per node)
KERNEL_BYTES = {"reset": (0, 16), "p2g": (100, 16), "grid": (0, 32), "g2p": (152, 16)}
# what the kernels of the hand-over pipeline actually have to move: P2G reads v, A, x
# (60 B); G2P reads x, F (, Jp) and writes every stream it owns.  Reported beside the SURVEY formula.
MOVED_BYTES = {"fixed_corotated": {"p2g": 60, "g2p": 48 + 96}, "snow": {"p2g": 60, "g2p": 52 + 100}}
E2E_SUBSTEPS_PER_SYNC = 20  # the reference main loop calls syncDevice every 20 advances (src/main.cu:99)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel, particles, nodes):
    """dram bytes read + written per launch of `kernel` from the committed ncu --set full capture of
    this workload (profiles/ncu_traffic.json, written from the capture named there); None if the capture
    was taken at another size."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        rec = json.load(open(p))[kernel]
        if rec["particles"] == particles and rec["nodes"] == nodes:
            return rec["dram_bytes_per_launch"]
    except Exception:
        pass
    return None


def workload(gpus, scaling="weak"):
    """Cubic N and balanced slabs for `gpus` ranks: weak = 2^26 particles per rank on a grid that
    grows with the rank count, strong = BASELINE.json configs[3] as is (N = 256, 2^26 particles in
    total) cut into thinner slabs."""
    from mpm_b200 import slabs

    if gpus == 1:
        return 256, [(0, 256)]
    N = 256 if scaling == "strong" else int(round(256 * gpus ** (1.0 / 3.0) / 2) * 2)
    return N, slabs.balanced_slabs(N, gpus, 0.1, 0.9)


class ClockSampler:
    """SM clock and throttle reasons during the timed region (B200_PROFILING.md recipe), sampled every 5 ms through
    NVML in a thread (a timed region of ~0.2 s falls between two lines of `nvidia-smi -lms`, whose queries take
    longer than that); nvidia-smi is the fallback when NVML cannot be loaded."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.index = index
        self.rows = []      # (t, sm_mhz, sm_max_mhz, [reason names])
        self.proc = None
        self.nvml = None
        self.running = False
        self.source = None

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[self.index])
            except Exception:
                pass
        return self.index

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index())
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.running = True
            self.source = "nvml"
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self._physical_index()), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi"
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        bits = {"hw_slowdown": getattr(n, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                "hw_thermal_slowdown": getattr(n, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": getattr(n, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                "sw_power_cap": getattr(n, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        while self.running:
            try:
                sm = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                mask = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                self.rows.append((time.monotonic(), sm, self.smax, [k for k, b in bits.items() if mask & b]))
            except Exception:
                pass
            time.sleep(0.005)

    def _read(self):
        for line in self.proc.stdout:
            c = [x.strip() for x in line.split(",")]
            try:
                self.rows.append((time.monotonic(), float(c[1]), float(c[2]),
                                  [n for n, v in zip(self.NAMES, c[5:9]) if v.lower().startswith("active")]))
            except Exception:
                pass

    def stop(self, t0=None, t1=None):
        """Samples taken inside [t0, t1] (host monotonic clock around the timed region); the sampler is
        started before the warm-up so that it is already streaming when the region begins."""
        if not self.nvml and not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no NVML, no nvidia-smi"]}
        time.sleep(0.05)
        self.running = False
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        inside = [r for r in self.rows if t0 is None or (t0 <= r[0] <= t1)]
        window = "timed region"
        if not inside:  # a very short region can fall between two samples: use the ones around it
            inside = [r for r in self.rows if t0 is None or (t0 - 0.5 <= r[0] <= t1 + 0.5)]
            window = "timed region +- 0.5 s (none fell inside)"
        sm = [r[1] for r in inside]
        reasons = sorted({n for r in inside for n in r[3]})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(r[2] for r in inside) if inside else None,
                "reasons": reasons, "samples": len(sm), "window": window, "source": self.source}


# --------------------------------------------------------------------------------------------------
# CPU arms (the only places that execute the checker under oracle/)
# --------------------------------------------------------------------------------------------------
def _cpu_threads():
    """All host cores: torch.distributed.run exports OMP_NUM_THREADS=1, which is not the machine."""
    import oracle_lib as ol

    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    ol.set_threads(n)
    return n


def _cpu_sample():
    import oracle_lib as ol
    import scenes

    N, P = 64, 1 << 20  # same generator, same 7.8 particles per cell as N=256 / 2^26
    p, mats = scenes.dense_block(P, N, density=P / 0.512)
    return N, P, p, mats, ol.new_grid(N)


def run_reference(args):
    """The reference has no CPU substep; this arm times the declared OpenMP transcription of its
    mpm.cu loops (oracle/, bit-exact against the reference's own headers) on all host cores, on a
    bounded sample of the same workload: the dense block at the same particles-per-cell."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle_lib as ol

    threads = _cpu_threads()
    N, P, p, mats, grid = _cpu_sample()
    for _ in range(args.warmup):
        ol.advance(p, mats, 1e-4, N, ol.FIXED_COROTATED, 1, grid)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ol.advance(p, mats, 1e-4, N, ol.FIXED_COROTATED, 1, grid)
    dt = time.perf_counter() - t0
    value = P * args.steps / dt
    sample = f"dense block N={N}, {P} particles (7.8 ppc as in the full workload), fixed-corotated, {args.steps} substeps, OpenMP {threads} threads"
    print(json.dumps({
        "impl": "reference", "metric": "particle_steps_per_s", "value": value, "unit": "particle-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "synthetic dense block, fixed-corotated (BASELINE.json configs[3]), bounded CPU sample",
                   "sample": sample},
        "cpu_baseline": {"value": value, "unit": "particle-steps/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def cpu_baseline(seconds_budget=20.0):
    import oracle_lib as ol

    threads = _cpu_threads()
    N, P, p, mats, grid = _cpu_sample()
    ol.advance(p, mats, 1e-4, N, ol.FIXED_COROTATED, 1, grid)
    steps, t0 = 0, time.perf_counter()
    while steps < 3 or (time.perf_counter() - t0 < seconds_budget / 2 and steps < 50):
        ol.advance(p, mats, 1e-4, N, ol.FIXED_COROTATED, 1, grid)
        steps += 1
    dt = time.perf_counter() - t0
    return {"value": P * steps / dt, "unit": "particle-steps/s", "cores": threads, "kind": "port",
            "sample": f"dense block N={N}, {P} particles (7.8 ppc), fixed-corotated, {steps} substeps, OpenMP {threads} threads"}


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
class Ctx:
    """Rank layout + helpers shared by the legs of one run."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist

        self.torch, self.dist, self.args = torch, dist, args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus:
            if self.rank == 0:
                print(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={self.world}; launch with torch.distributed.run", file=sys.stderr)
            sys.exit(2)
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the substep has no CPU fallback")
        torch.cuda.set_device(self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, values, op="sum"):
        t = self.torch.tensor(values, dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM if op == "sum" else self.dist.ReduceOp.MAX)
        return [float(v) for v in t]

    def fail(self, what):
        print(f"bench.py: self-check failed on rank {self.rank}: {what}", file=sys.stderr, flush=True)
        sys.exit(3)


class numa_local:
    """Runs its body on the CPUs closest to this rank's GPU (NVML's affinity mask), so that the pinned
    host buffer allocated inside is placed on that NUMA node: 8 ranks x 2 x 7 GB per frame otherwise
    cross the socket interconnect.  No-op when NVML is not available."""

    def __init__(self, device_index):
        self.device_index = device_index
        self.saved = None

    def __enter__(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.device_index)
            n_cpus = os.cpu_count() or 1
            words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpus + 63) // 64)
            cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
            allowed = os.sched_getaffinity(0)
            if cpus & allowed:
                self.saved = allowed
                os.sched_setaffinity(0, cpus & allowed)
        except Exception:
            self.saved = None
        return self

    def __exit__(self, *exc):
        if self.saved is not None:
            os.sched_setaffinity(0, self.saved)
        return False


def make_sim(ctx, model, scaling, particles_per_gpu, svd, sort_every, pipeline="handover", p2g="runs", g2p="tile", rebin_permille=0, ghost=0):
    import mpm_b200

    emu = getattr(ctx.args, "emulate_rank", None)
    if emu:  # experiments: the slab, grid and particle share of rank R of W on ONE GPU, without neighbours
        r, w = (int(v) for v in emu.split("/"))
        N, slabs = workload(w, scaling)
        xb, xe = slabs[r]
        P_total = particles_per_gpu * (w if scaling == "weak" else 1)
        slabs = [slabs[r]]
    else:
        N, slabs = workload(ctx.world, scaling)
        xb, xe = slabs[ctx.rank]
        P_total = particles_per_gpu * (ctx.world if scaling == "weak" else 1)
    density = P_total / 0.512  # --particle-count: the block fills 0.8^3 of the unit cube
    snow = model == "snow"
    if snow:  # scenes/snowman.toml material
        mats = mpm_b200.make_material(1.0 / density, 700.0, 1.4e5, 0.2, 10.0, 0.975, 1.0075)
    else:
        mats = mpm_b200.make_material(1.0 / density, 1000.0, 1.4e5, 0.2, 0.0, 0.0, 1e30)
    cap = int(P_total / ctx.world * 1.15) if ctx.world > 1 else (int(particles_per_gpu * 1.15) if emu else 0)
    sim = mpm_b200.Sim(N, 1e-4, mats, model=mpm_b200.SNOW if snow else mpm_b200.FIXED_COROTATED,
                       svd_mode=mpm_b200.SVD_FAST if svd == "fast" else mpm_b200.SVD_EXACT, sort_every=sort_every,
                       x_begin=xb, x_end=xe, device=ctx.local_rank, capacity=cap,
                       p2g_mode=mpm_b200.P2G_RUNS if p2g == "runs" else mpm_b200.P2G_DIRECT,
                       g2p_mode=mpm_b200.G2P_TILE if g2p == "tile" else mpm_b200.G2P_DIRECT,
                       pipeline=mpm_b200.PIPE_HANDOVER if pipeline == "handover" else mpm_b200.PIPE_CLASSIC,
                       rebin_permille=rebin_permille, ghost=ghost)
    if ctx.world > 1:
        from mpm_b200 import slabs as _slabs

        sim.attach_comm(_slabs.share_unique_id(ctx.dist, ctx.rank, mpm_b200.comm_unique_id), ctx.rank, ctx.world)
    return sim, N, slabs, P_total, float(mats[1])


def timed(ctx, sim, steps, warmup):
    """W untimed substeps, then K substeps between CUDA events on the handle's stream, barrier +
    synchronize on both sides; returns (ms max over ranks, launches, host monotonic window)."""
    torch = ctx.torch
    stream = torch.cuda.ExternalStream(sim.stream)
    left = warmup
    while left > 0:  # in calls of several substeps, so that the hand-over pipeline is warm too
        sim.advance(min(left, 4))
        left -= 4
    sim.sync()
    ctx.barrier()
    t0 = time.monotonic()
    launches0 = sim.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    sim.advance(steps)
    e1.record(stream)
    sim.sync()
    ctx.barrier()
    ms = ctx.reduce([e0.elapsed_time(e1)], "max")[0]
    return ms, sim.launches - launches0, (t0, time.monotonic())


def short_leg(ctx, model, scaling, svd, steps, shear=0.0, f_noise=0.0, particles=P_PER_GPU, sort_every=8, rebin_permille=0):
    """One more workload, measured like the main one but shorter; returns a small dict."""
    ghost = 0
    if shear and ctx.world > 1:
        # slab handles: the fastest particles (0.4 shear m/s at the block faces) must stay within the ghost
        # planes between two re-bins, or the re-bin fails (MpmDiagnostics.escaped): 2 ghost planes, cadence 4
        ghost, sort_every = 2, 4
    sim, N, slabs, P_total, _ = make_sim(ctx, model, scaling, particles, svd, sort_every, ghost=ghost, rebin_permille=rebin_permille)
    sim.generate_dense_block(P_total, seed=1234, shear=shear, f_noise=f_noise)
    sim.sync()
    rebins0 = sim.rebins
    ms, _, _ = timed(ctx, sim, steps, 8)
    rebins = sim.rebins - rebins0
    P_all = ctx.reduce([float(sim.count)])[0]
    diag = sim.diagnostics()
    sim.close()
    if abs(P_all - P_total) > 0.5:
        ctx.fail(f"{model}/{scaling}: {P_all} particles on the ranks, {P_total} generated")
    return {"N": N, "particles": int(P_all), "model": model, "svd_mode": svd, "scaling": scaling if ctx.world > 1 else "n/a",
            "ms_per_step": ms / steps, "value": P_all / (ms / steps * 1e-3), "unit": "particle-steps/s", "steps": steps,
            "shear": shear, "f_noise": f_noise, "sort_every": sort_every, "rebin_permille": rebin_permille, "rebins": int(rebins),
            "ghost": ghost, "nonfinite": diag["nonfinite"]}


def parity_nranks(ctx):
    """A small sheared block on the SAME number of slabs, every rank holding and exchanging particles
    in both directions, against the CPU checker on rank 0 (per-particle position / velocity error)."""
    import mpm_b200
    import oracle_lib as ol
    import scenes
    from mpm_b200 import slabs as sl

    W = ctx.world
    N = max(32, 8 * W)  # slabs of >= 8 planes
    P, steps, dt = 200_000, 40, 1e-4
    # v_x = shear * (y - 0.5): +x above the mid-plane, -x below, so particles cross every slab boundary both ways
    p, mats = scenes.dense_block(P, N, kind=ol.SNOW, shear=40.0, f_noise=0.01)
    slabs = sl.balanced_slabs(N, W, 0.1, 0.9)
    own = sl.owner(p["x"][:, 0], N, slabs)
    xb, xe = slabs[ctx.rank]
    sim = mpm_b200.Sim(N, dt, mats, model=mpm_b200.SNOW, svd_mode=mpm_b200.SVD_EXACT, sort_every=4, x_begin=xb, x_end=xe,
                       device=ctx.local_rank, capacity=P)
    sim.attach_comm(sl.share_unique_id(ctx.dist, ctx.rank, mpm_b200.comm_unique_id), ctx.rank, W)
    mine = np.where(own == ctx.rank)[0]
    sim.upload_with_ids(np.ascontiguousarray(p[mine]), mine.astype(np.uint32))
    sim.advance(steps)
    got = sim.download()
    _, ids = sim.sort_state()
    diag = sim.diagnostics()
    sim.close()
    box = [None] * W if ctx.rank == 0 else None
    ctx.dist.gather_object((ids, got, len(mine), diag), box, dst=0)
    out = None
    if ctx.rank == 0:
        _cpu_threads()
        ref, _ = ol.advance(p.copy(), mats, dt, N, ol.SNOW, steps)
        merged = np.zeros_like(p)
        seen = np.zeros(P, np.int32)
        moved_up = moved_down = 0
        for r, (ids_r, got_r, n0, _) in enumerate(box):
            merged[ids_r] = got_r
            seen[ids_r] += 1
            moved_up += int((own[ids_r] < r).sum())
            moved_down += int((own[ids_r] > r).sum())
        ok_once = bool((seen == 1).all())
        err_x = float(np.abs(merged["x"].astype(np.float64) - ref["x"]).max() * N) if ok_once else float("inf")
        err_v = float(np.abs(merged["v"].astype(np.float64) - ref["v"]).max()) if ok_once else float("inf")
        vmax = float(np.abs(ref["v"]).max())
        out = {"ranks": W, "N": N, "particles": P, "substeps": steps, "model": "snow", "svd_mode": "exact",
               "max_dx_over_dx": err_x, "max_dv": err_v, "v_max": vmax, "migrated_up": moved_up, "migrated_down": moved_down,
               "escaped": int(sum(b[3]["escaped"] for b in box)), "every_particle_on_one_rank": ok_once,
               "tolerance": {"dx_over_dx": 1e-3, "dv_over_vmax": 1e-3},
               "ok": ok_once and err_x < 1e-3 and err_v < 1e-3 * vmax and moved_up > 0 and moved_down > 0}
    flag = ctx.reduce([0.0 if (out is None or out["ok"]) else 1.0])[0]
    if flag:
        ctx.fail(f"parity on {W} slabs: {out}")
    return out


def small_scenes():
    """BASELINE.json configs[0..2] (the reference's README command lines, stand-in meshes) through the
    host facade: ms per substep, with the launches replayed as CUDA graphs (default for small scenes)
    and without.  These scenes are launch-bound, not bandwidth-bound."""
    from mpm_b200 import host

    out = {}
    cwd = os.getcwd()
    os.chdir(ROOT)
    try:
        for name, line in (("config0_rubber_duck", "--scene scenes/rubber_duck.toml -N 16 --particle-count 10000"),
                           ("config1_snowman", "--scene scenes/snowman.toml -N 64 --particle-count 500000"),
                           ("config2_liquid_bunny", "--scene scenes/liquid_bunny.toml -N 32 --particle-count 1000000")):
            rec = {"command": line}
            for graphs in ("auto", "off"):
                s = host.Scene(*line.split(), "--graphs", graphs, "--svd", "fast")
                n = len(s.active_particles())
                s.init_cuda()
                s.advance(200)
                s.sync_device()
                steps = 2000
                t0 = time.perf_counter()
                s.advance(steps)
                s.sync_device()   # includes one download of the particles, like the reference's loop every 20 substeps
                dt = time.perf_counter() - t0
                rec["particles"] = n
                rec["ms_per_substep_graphs_" + graphs] = 1e3 * dt / steps
                s.close()
            out[name] = rec
    finally:
        os.chdir(cwd)
    return out


def invariants(ctx, sim, P_total, mass, slab, N):
    """Size-independent self-checks on the benchmark state (outside every timed region)."""
    P_all = ctx.reduce([float(sim.count)])[0]
    if abs(P_all - P_total) > 0.5:
        ctx.fail(f"{P_all} particles on the ranks, {P_total} generated")
    sim.stage("reset_grid")
    sim.stage("p2g")  # on slab handles including the halo exchange
    g = sim.grid()
    x0 = max(0, slab[0] - 1) if ctx.world > 1 else 0  # first local plane (ghost = 1)
    own = g[slab[0] - x0: slab[1] - x0]
    finite = bool(np.isfinite(g).all())
    m_all = ctx.reduce([float(own[..., 3].sum(dtype=np.float64))])[0]
    del g, own
    x = sim.download_positions()
    finite = finite and bool(np.isfinite(x).all())
    inside = bool(((x > 0.0) & (x < 1.0)).all())
    del x
    diag = sim.diagnostics()
    bad = ctx.reduce([0.0 if finite and inside else 1.0, float(diag["escaped"]), float(diag["nonfinite"])])
    rel = abs(m_all - P_total * mass) / (P_total * mass)
    out = {"particles": int(P_all), "particles_expected": int(P_total), "grid_mass_rel_err": rel, "grid_mass_tolerance": 1e-5,
           "finite_and_inside_domain": bad[0] == 0.0, "escaped": int(bad[1]), "nonfinite": int(bad[2]),
           "ok": bad[0] == 0.0 and bad[1] == 0.0 and bad[2] == 0.0 and rel < 1e-5}
    if not out["ok"]:
        ctx.fail(f"invariants: {out}")
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--sort-every", type=int, default=8)
    ap.add_argument("--svd", default="fast", choices=["fast", "exact"])
    ap.add_argument("--particles", type=int, default=P_PER_GPU, help="particles per GPU (default 2^26)")
    ap.add_argument("--g2p", default="tile", choices=["tile", "direct"], help="G2P kernel (MpmParams.g2p_mode)")
    ap.add_argument("--p2g", default="runs", choices=["runs", "direct"], help="P2G kernel (MpmParams.p2g_mode)")
    ap.add_argument("--pipeline", default="handover", choices=["handover", "classic"], help="substep pipeline (MpmParams.pipeline)")
    ap.add_argument("--shear", type=float, default=0.0, help="stressed block: shear velocity field [1/s] (mpm_generate_dense_block_stressed)")
    ap.add_argument("--f-noise", type=float, default=0.0, help="stressed block: amplitude of the perturbation of F")
    ap.add_argument("--rebin-permille", type=int, default=0,
                    help="MpmParams.rebin_permille: also re-bin on measured disorder (0 = fixed cadence only, the default)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N > 1: weak = 2^26 particles per GPU (default), strong = 2^26 particles in total at N = 256")
    ap.add_argument("--model", default="fixed_corotated", choices=["fixed_corotated", "snow"],
                    help="material model; snow = BASELINE.json configs[4] (plasticity via svd3 in G2P)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--emulate-rank", default=None, metavar="R/W",
                    help="experiments on one GPU: the slab geometry (N, planes, particle share) of rank R of W, without neighbours; "
                         "use with --no-extras --no-e2e (the invariants are not checked: the slab has no neighbours to exchange with)")
    ap.add_argument("--no-extras", action="store_true", help="skip the extra configurations / stressed workloads / N-rank parity")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    ctx = Ctx(args)
    torch, rank, world = ctx.torch, ctx.rank, ctx.world
    full = args.particles == P_PER_GPU and not args.no_extras

    parity = parity_nranks(ctx) if (world > 1 and not args.no_extras) else None

    sim, N, slabs, P_total, mass = make_sim(ctx, args.model, args.scaling, args.particles, args.svd, args.sort_every, args.pipeline,
                                            args.p2g, args.g2p, args.rebin_permille)
    sim.generate_dense_block(P_total, seed=1234, shear=args.shear, f_noise=args.f_noise)
    sim.sync()
    P_local, G_local = sim.count, sim.grid_nodes
    inv = None if args.emulate_rank else invariants(ctx, sim, P_total, mass, slabs[rank], N)

    sampler = ClockSampler(ctx.local_rank)
    if rank == 0:
        sampler.start()
    ms, launches, window = timed(ctx, sim, args.steps, args.warmup)
    clocks = sampler.stop(*window) if rank == 0 else None
    P_all, G_all = ctx.reduce([float(P_local), float(G_local)])
    ms_per_step = ms / args.steps
    value = P_all / (ms_per_step * 1e-3)

    # per-stage times (events + sync around every stage: serialised, used only for shares and the
    # dominant kernel's own duration)
    sim.stage_times()
    n_prof = min(args.steps, 16)
    sim.advance(n_prof)
    st = sim.stage_times()
    sim.set_stage_timing(False)
    ctx.barrier()
    peak, peak_src = peaks()
    substep_keys = ("reset", "p2g", "grid", "g2p")
    dom = max(substep_keys, key=lambda s: st[s])
    bp, bn = KERNEL_BYTES[dom]
    dom_ms = st[dom] / n_prof
    achieved = (bp * P_local + bn * G_local) / (dom_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic(dom, P_local, G_local), "peak_source": peak_src, "kernel_ms": dom_ms,
                "algorithmic_bytes_per_launch": bp * P_local + bn * G_local}
    per_kernel = {}
    for s in substep_keys:
        kp, kn = KERNEL_BYTES[s]
        t_ms = st[s] / n_prof
        if t_ms <= 0:
            continue
        rec = {"ms": t_ms, "frac": (kp * P_local + kn * G_local) / (t_ms * 1e-3) / 1e9 / peak}
        if args.pipeline == "handover" and s in MOVED_BYTES[args.model]:
            rec["frac_on_bytes_moved"] = (MOVED_BYTES[args.model][s] * P_local + kn * G_local) / (t_ms * 1e-3) / 1e9 / peak
        per_kernel[s] = rec
    sub_ach = (BYTES_PER_PARTICLE * P_all + BYTES_PER_NODE * G_all) / (ms_per_step * 1e-3) / 1e9 / world
    substep_roofline = {"achieved_per_gpu": sub_ach, "peak": peak, "unit": "GB/s", "frac": sub_ach / peak,
                        "bytes": "252*P + 80*G per substep (SURVEY.md 8(d))", "per_kernel": per_kernel,
                        "note": "frac = SURVEY 8(d) algorithmic bytes / time / measured copy peak; with the hand-over pipeline the kernels "
                                "move fewer bytes than that (P2G 60 B/particle instead of 100): per_kernel.frac_on_bytes_moved"}
    stage_ms = {k: v / n_prof for k, v in st.items()}

    # e2e: host AoS buffers through the C ABI, the reference's own cadence: upload (initCuda),
    # 20 x advance, syncDevice (download all particles) — copies inside the timed region.
    e2e = None
    if not args.no_e2e:
        n_host = P_local
        with numa_local(ctx.local_rank):
            host = torch.empty(n_host * 104, dtype=torch.uint8, pin_memory=True)
            host.zero_()
            host_out = torch.empty(n_host * 104, dtype=torch.uint8, pin_memory=True)
            host_out.zero_()
        got = sim.download_ptr(host.data_ptr(), n_host)  # current state as the host-side truth
        assert got == n_host
        ctx.barrier()
        # (a) strictly in turn, as the reference's loop does it: upload, 20 substeps, blocking download
        frames = 2
        t0 = time.perf_counter()
        for _ in range(frames):
            sim.upload_ptr(host.data_ptr(), n_host)
            sim.advance(E2E_SUBSTEPS_PER_SYNC)
            sim.download_ptr(host_out.data_ptr(), n_host)
        ctx.barrier()
        wall_seq = ctx.reduce([time.perf_counter() - t0], "max")[0]
        # (b) the same frames with the PCIe copies overlapped (mpm_prefetch_particles_aos,
        # mpm_download_particles_aos_async): frame k + 1's input travels to the device and frame k's result to
        # the host while the substeps of a frame run.  Every frame still uploads its 104 B/particle input from
        # pinned host memory and reads its full result back, inside the timed region.
        frames_p = 4
        sim.prefetch_ptr(host.data_ptr(), n_host)
        sim.upload_ptr(host.data_ptr(), n_host)          # one untimed frame fills the pipeline
        sim.prefetch_ptr(host.data_ptr(), n_host)
        sim.advance(E2E_SUBSTEPS_PER_SYNC)
        sim.download_ptr_async(host_out.data_ptr(), n_host)
        sim.sync()
        ctx.barrier()
        t0 = time.perf_counter()
        for _ in range(frames_p):
            sim.upload_ptr(host.data_ptr(), n_host)      # the copy announced one frame ago
            sim.prefetch_ptr(host.data_ptr(), n_host)
            sim.advance(E2E_SUBSTEPS_PER_SYNC)
            sim.download_ptr_async(host_out.data_ptr(), n_host)
        sim.download_wait()
        sim.sync()
        ctx.barrier()
        wall = ctx.reduce([time.perf_counter() - t0], "max")[0]
        e2e = {"value": P_all * E2E_SUBSTEPS_PER_SYNC * frames_p / wall, "unit": "particle-steps/s",
               "h2d_bytes_per_step": n_host * 104 / E2E_SUBSTEPS_PER_SYNC, "d2h_bytes_per_step": n_host * 104 / E2E_SUBSTEPS_PER_SYNC,
               "substeps_per_sync": E2E_SUBSTEPS_PER_SYNC, "frames": frames_p,
               "sequential_value": P_all * E2E_SUBSTEPS_PER_SYNC * frames / wall_seq,
               "what": "per frame: mpm_upload_particles_aos (pinned host AoS, 104 B/particle) + 20 x mpm_advance + full download; "
                       "value: copies overlapped with the substeps of the neighbouring frames (mpm_prefetch_particles_aos / "
                       "mpm_download_particles_aos_async, one H2D and one D2H of all particles per frame inside the timed region); "
                       "sequential_value: the same calls strictly in turn (blocking upload and download); bytes are per substep"}
        del host_out
        del host
    sim.close()

    # the other BASELINE.json configurations and stressed variants of the workload, each its own short
    # run (2^26 particles per GPU, N as the main run): driver-visible numbers for configs[3] strong
    # scaling and configs[4] (snow), and for inputs that are not at rest (the quiescent block
    # is the cheapest case of every data-dependent path)
    configs, stressed = None, None
    if full:
        k = max(8, min(args.steps, 16))
        configs = {"config5_snow": short_leg(ctx, "snow", args.scaling, "fast", k)}
        if world > 1:
            configs["config4_fixed_corotated_strong"] = short_leg(ctx, "fixed_corotated", "strong", "fast", k)
        stressed = {
            "what": "same block, v = shear * (y - 0.5, 0, 0.3 (x - 0.5)) (0.2 cells per substep at the faces for shear 20 at N = 256) and "
                    "F = I + f_noise * u[-1,1): Newton polar iterates, snow yields (|F - I| up to 5 % against an elastic range of "
                    "-2.5 % / +0.75 %), particles cross cells between re-bins",
            "fixed_corotated_sheared": short_leg(ctx, "fixed_corotated", args.scaling, "fast", k, shear=20.0, f_noise=0.05),
            "snow_yielding": short_leg(ctx, "snow", args.scaling, "fast", k, shear=20.0, f_noise=0.05),
        }
        if world == 1:
            stressed["snow_exact_svd"] = short_leg(ctx, "snow", args.scaling, "exact", 8)
            # re-bin on measured disorder instead of a fixed cadence (MpmParams.rebin_permille: when the cell crossings
            # G2P counts since the last re-bin pass 5 % of the particles); `rebins` = re-bins in warm-up + timed substeps
            configs["adaptive_rebin_at_rest"] = short_leg(ctx, "fixed_corotated", args.scaling, "fast", k, sort_every=0, rebin_permille=50)
            stressed["adaptive_rebin_sheared"] = short_leg(ctx, "fixed_corotated", args.scaling, "fast", k, shear=20.0, f_noise=0.05,
                                                           sort_every=0, rebin_permille=50)
    small = small_scenes() if (full and world == 1) else None

    if rank == 0:
        snow = args.model == "snow"
        out = {
            "metric": "particle_steps_per_s", "value": value, "unit": "particle-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": args.scaling if world > 1 else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"synthetic dense block N={N}, {int(P_all)} particles, {args.model} (BASELINE.json configs[{4 if snow else 3}]"
                                   + (")" if world == 1 else (f" scaled weakly to {world} GPUs: 2^26 particles and ~2^24 nodes per GPU)" if args.scaling == "weak"
                                                              else f" strong scaling: the same block cut into {world} slabs)")),
                       "N": N, "particles": int(P_all), "grid_nodes": int(G_all), "dt": 1e-4, "model": args.model,
                       "svd_mode": args.svd, "sort_every": args.sort_every, "rebin_permille": args.rebin_permille, "p2g": args.p2g, "g2p": args.g2p,
                       "pipeline": args.pipeline, "shear": args.shear, "f_noise": args.f_noise, "slabs": slabs if world > 1 else None,
                       "l2": "inputs (6.7 GB particles + 268 MB grid per GPU) are far larger than the 126 MB L2; no flush"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
            "substep_roofline": substep_roofline, "stage_ms": stage_ms, "invariants": inv,
        }
        if parity is not None:
            out["parity_nranks"] = parity
        if configs is not None:
            out["configs"] = configs
        if stressed is not None:
            out["stressed"] = stressed
        if small is not None:
            out["small_scenes"] = small
        if world == 1 and not args.no_cpu:
            out["cpu_baseline"] = cpu_baseline()
        print(json.dumps(out))
    if world > 1:
        ctx.dist.destroy_process_group()


if __name__ == "__main__":
    main()
