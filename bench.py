#!/usr/bin/env python
"""Benchmark of the MLS-MPM substep (BASELINE.json metric: particle-steps/s + HBM roofline).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N=1 workload = BASELINE.json configs[3]: synthetic dense block, N=256, 2^26 particles,
fixed-corotated.  For N>1 (launched with torch.distributed.run, one
rank per GPU) every rank keeps 2^26 particles and ~2^24 grid nodes (weak scaling): the cubic
domain grows to N = 256 * gpus^(1/3) and is cut into x-slabs balanced by particle count.
A "step" is one substep (grid reset -> P2G -> [halo exchange] -> grid update -> G2P), the sort is
included at its cadence (--sort-every).  Inputs are generated on the device and are far larger
than L2 (6.7 GB particles + 268 MB grid per GPU), so no L2 flush is needed between steps.
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
# (60 B), G2P reads x, F (, Jp) and writes 24 (25) streams; reported beside the SURVEY formula
MOVED_BYTES = {"reset": (0, 16), "p2g": (60, 16), "grid": (0, 32), "g2p": (144, 16)}
E2E_SUBSTEPS_PER_SYNC = 20  # the reference main loop calls syncDevice every 20 advances (src/main.cu:99)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel, particles, nodes):
    """dram bytes read + written per launch of `kernel` from the committed ncu --set full capture of
    this workload (profiles/ncu_traffic.json, written by tools/ncu_summary.py); None if the capture
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
    """nvidia-smi clocks during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.monotonic(), [c.strip() for c in line.split(",")]))

    def stop(self, t0=None, t1=None):
        """Samples taken inside [t0, t1] (host monotonic clock around the timed region); the sampler is
        started before the warm-up so that nvidia-smi is already streaming when the region begins."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        inside = [r for (t, r) in self.rows if t0 is None or (t0 <= t <= t1 + 0.03)]
        window = "timed region"
        if not inside:  # a very short region can fall between two samples: use the ones around it
            inside = [r for (t, r) in self.rows if t0 is None or (t0 - 0.5 <= t <= t1 + 0.5)]
            window = "timed region +- 0.5 s (none fell inside)"
        for r in inside:
            try:
                sm.append(float(r[1]))
                smax.append(float(r[2]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


def run_reference(args):
    """The reference has no CPU substep; this arm times the declared OpenMP transcription of its
    mpm.cu loops (oracle/, bit-exact against the reference's own headers) on the host cores, on a
    bounded sample of the same workload: the dense block at the same particles-per-cell."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle_lib as ol
    import scenes

    N, P = 64, 1 << 20  # same generator, same 7.8 particles per cell as N=256 / 2^26
    p, mats = scenes.dense_block(P, N, density=P / 0.512)
    grid = ol.new_grid(N)
    threads = ol.max_threads()
    for _ in range(args.warmup):
        ol.advance(p, mats, 1e-4, N, ol.FIXED_COROTATED, 1, grid)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ol.advance(p, mats, 1e-4, N, ol.FIXED_COROTATED, 1, grid)
    dt = time.perf_counter() - t0
    value = P * args.steps / dt
    sample = f"dense block N={N}, {P} particles (7.8 ppc as in the full workload), fixed-corotated, {args.steps} substeps"
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
    import scenes

    N, P = 64, 1 << 20
    p, mats = scenes.dense_block(P, N, density=P / 0.512)
    grid = ol.new_grid(N)
    threads = ol.max_threads()
    ol.advance(p, mats, 1e-4, N, ol.FIXED_COROTATED, 1, grid)
    steps, t0 = 0, time.perf_counter()
    while steps < 3 or (time.perf_counter() - t0 < seconds_budget / 2 and steps < 50):
        ol.advance(p, mats, 1e-4, N, ol.FIXED_COROTATED, 1, grid)
        steps += 1
    dt = time.perf_counter() - t0
    return {"value": P * steps / dt, "unit": "particle-steps/s", "cores": threads, "kind": "port",
            "sample": f"dense block N={N}, {P} particles (7.8 ppc), fixed-corotated, {steps} substeps, OpenMP {threads} threads"}


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
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    import mpm_b200

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if rank == 0:
            print(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={world}; launch with torch.distributed.run", file=sys.stderr)
        sys.exit(2)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the substep has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    N, slabs = workload(world, args.scaling)
    xb, xe = slabs[rank]
    P_total = args.particles * (world if args.scaling == "weak" else 1)
    dt = 1e-4
    density = P_total / 0.512  # --particle-count: the block fills 0.8^3 of the unit cube
    snow = args.model == "snow"
    if snow:  # scenes/snowman.toml material
        mats = mpm_b200.make_material(1.0 / density, 700.0, 1.4e5, 0.2, 10.0, 0.975, 1.0075)
    else:
        mats = mpm_b200.make_material(1.0 / density, 1000.0, 1.4e5, 0.2, 0.0, 0.0, 1e30)
    svd_mode = mpm_b200.SVD_FAST if args.svd == "fast" else mpm_b200.SVD_EXACT
    cap = int(P_total / world * 1.15) if world > 1 else 0
    sim = mpm_b200.Sim(N, dt, mats, model=mpm_b200.SNOW if snow else mpm_b200.FIXED_COROTATED, svd_mode=svd_mode, sort_every=args.sort_every,
                       x_begin=xb, x_end=xe, device=local_rank, capacity=cap,
                       p2g_mode=mpm_b200.P2G_RUNS if args.p2g == "runs" else mpm_b200.P2G_DIRECT,
                       g2p_mode=mpm_b200.G2P_TILE if args.g2p == "tile" else mpm_b200.G2P_DIRECT,
                       pipeline=mpm_b200.PIPE_HANDOVER if args.pipeline == "handover" else mpm_b200.PIPE_CLASSIC,
                       rebin_permille=args.rebin_permille)
    if world > 1:
        from mpm_b200 import slabs as _slabs

        sim.attach_comm(_slabs.share_unique_id(dist, rank, mpm_b200.comm_unique_id), rank, world)
    sim.generate_dense_block(P_total, seed=1234, shear=args.shear, f_noise=args.f_noise)
    sim.sync()
    P_local = sim.count
    G_local = sim.grid_nodes

    stream = torch.cuda.ExternalStream(sim.stream)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        sim.advance(1)
    sim.sync()
    barrier()
    t_region0 = time.monotonic()
    launches0 = sim.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    sim.advance(args.steps)
    e1.record(stream)
    sim.sync()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = sim.launches - launches0
    clocks = sampler.stop(t_region0, time.monotonic()) if rank == 0 else None
    t = torch.tensor([ms, float(P_local), float(G_local)], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms = float(tmax[0])
        P_all, G_all = float(tsum[1]), float(tsum[2])
    else:
        P_all, G_all = float(P_local), float(G_local)
    ms_per_step = ms / args.steps
    value = P_all / (ms_per_step * 1e-3)

    # per-stage times (events + sync around every stage: serialised, used only for shares and the
    # dominant kernel's own duration)
    sim.stage_times()
    n_prof = min(args.steps, 16)
    sim.advance(n_prof)
    st = sim.stage_times()
    sim.set_stage_timing(False)
    barrier()
    peak, peak_src = peaks()
    substep_keys = ("reset", "p2g", "grid", "g2p")
    dom = max(substep_keys, key=lambda s: st[s])
    bp, bn = KERNEL_BYTES[dom]
    dom_ms = st[dom] / n_prof
    achieved = (bp * P_local + bn * G_local) / (dom_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic(dom, P_local, G_local), "peak_source": peak_src, "kernel_ms": dom_ms,
                "algorithmic_bytes_per_launch": bp * P_local + bn * G_local}
    sub_ach = (BYTES_PER_PARTICLE * P_all + BYTES_PER_NODE * G_all) / (ms_per_step * 1e-3) / 1e9 / world
    substep_roofline = {"achieved_per_gpu": sub_ach, "peak": peak, "unit": "GB/s", "frac": sub_ach / peak,
                        "bytes": "252*P + 80*G per substep (BASELINE.md)"}
    stage_ms = {k: v / n_prof for k, v in st.items()}

    # e2e: host AoS buffers through the C ABI, the reference's own cadence: upload (initCuda),
    # 20 x advance, syncDevice (download all particles) — copies inside the timed region.
    e2e = None
    if not args.no_e2e:
        n_host = P_local
        host = torch.empty(n_host * 104, dtype=torch.uint8, pin_memory=True)
        got = sim.download_ptr(host.data_ptr(), n_host)  # current state as the host-side truth
        assert got == n_host
        barrier()
        frames = 2
        t0 = time.perf_counter()
        for _ in range(frames):
            sim.upload_ptr(host.data_ptr(), n_host)
            sim.advance(E2E_SUBSTEPS_PER_SYNC)
            sim.download_ptr(host.data_ptr(), n_host)
        barrier()
        wall = time.perf_counter() - t0
        tw = torch.tensor([wall], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tw, op=dist.ReduceOp.MAX)
        wall = float(tw[0])
        e2e = {"value": P_all * E2E_SUBSTEPS_PER_SYNC * frames / wall, "unit": "particle-steps/s",
               "h2d_bytes_per_step": n_host * 104 / E2E_SUBSTEPS_PER_SYNC, "d2h_bytes_per_step": n_host * 104 / E2E_SUBSTEPS_PER_SYNC,
               "substeps_per_sync": E2E_SUBSTEPS_PER_SYNC,
               "what": "mpm_upload_particles_aos (pinned host AoS, 104 B/particle) + 20 x mpm_advance + mpm_download_particles_aos per frame; bytes are per substep"}
        del host

    if rank == 0:
        out = {
            "metric": "particle_steps_per_s", "value": value, "unit": "particle-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": args.scaling if world > 1 else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"synthetic dense block N={N}, {int(P_all)} particles, {args.model} (BASELINE.json configs[{4 if snow else 3}]"
                                   + (")" if world == 1 else (f" scaled weakly to {world} GPUs: 2^26 particles and ~2^24 nodes per GPU)" if args.scaling == "weak"
                                                              else f" strong scaling: the same block cut into {world} slabs)")),
                       "N": N, "particles": int(P_all), "grid_nodes": int(G_all), "dt": dt, "model": args.model,
                       "svd_mode": args.svd, "sort_every": args.sort_every, "rebin_permille": args.rebin_permille, "p2g": args.p2g, "g2p": args.g2p, "pipeline": args.pipeline, "shear": args.shear, "f_noise": args.f_noise, "slabs": slabs if world > 1 else None,
                       "l2": "inputs (6.7 GB particles + 268 MB grid per GPU) are far larger than the 126 MB L2; no flush"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
            "substep_roofline": substep_roofline, "stage_ms": stage_ms,
        }
        if world == 1 and not args.no_cpu:
            out["cpu_baseline"] = cpu_baseline()
        print(json.dumps(out))
    sim.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
