"""Where does the pipelined e2e frame spend its time?  python tools/e2e_probe.py  (GPU box)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mpm_b200

P, N = 1 << 26, 256
mats = mpm_b200.make_material(0.512 / P, 1000.0, 1.4e5, 0.2, 0.0, 0.0, 1e30)
sim = mpm_b200.Sim(N, 1e-4, mats, model=mpm_b200.FIXED_COROTATED, svd_mode=mpm_b200.SVD_FAST, sort_every=8, device=0)
sim.generate_dense_block(P, seed=1234)
sim.advance(8); sim.sync()
host = torch.empty(P * 104, dtype=torch.uint8, pin_memory=True); host.zero_()
out = torch.empty(P * 104, dtype=torch.uint8, pin_memory=True); out.zero_()
sim.download_ptr(host.data_ptr(), P)

def frames(n, up=True, pre=True, adv=True, down="async"):
    if pre and up:
        sim.prefetch_ptr(host.data_ptr(), P)
    sim.sync(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(n):
        if up:
            sim.upload_ptr(host.data_ptr(), P)
            if pre:
                sim.prefetch_ptr(host.data_ptr(), P)
        if adv:
            sim.advance(20)
        if down == "async":
            sim.download_ptr_async(out.data_ptr(), P)
        elif down == "block":
            sim.download_ptr(out.data_ptr(), P)
    sim.download_wait(); sim.sync(); torch.cuda.synchronize()
    return (time.perf_counter() - t) / n * 1e3

print("advance(20) only                      %.1f ms/frame" % frames(4, up=False, down=None))
print("blocking upload only                  %.1f" % frames(4, pre=False, adv=False, down=None))
print("prefetched upload only                %.1f" % frames(6, adv=False, down=None))
print("blocking download only                %.1f" % frames(4, up=False, adv=False, down="block"))
print("async download only                   %.1f" % frames(6, up=False, adv=False, down="async"))
print("prefetched upload + advance           %.1f" % frames(6, down=None))
print("advance + async download              %.1f" % frames(6, up=False))
print("prefetched upload + async download    %.1f" % frames(6, adv=False))
print("all, pipelined                        %.1f" % frames(8))
print("all, blocking                         %.1f" % frames(3, pre=False, down="block"))
