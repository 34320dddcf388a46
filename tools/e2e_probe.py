import sys, time, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import torch, numpy as np, mpm_b200
P = 1 << 26
mats = mpm_b200.make_material(0.512 / P, 1000.0, 1.4e5, 0.2, 0.0, 0.0, 1e30)
sim = mpm_b200.Sim(256, 1e-4, mats, model=mpm_b200.FIXED_COROTATED, svd_mode=mpm_b200.SVD_FAST, sort_every=8)
sim.generate_dense_block(P); sim.sync()
host = torch.empty(P * 104, dtype=torch.uint8, pin_memory=True)
sim.download_ptr(host.data_ptr(), P)
for f in range(3):
    t0 = time.perf_counter(); sim.upload_ptr(host.data_ptr(), P); sim.sync(); t1 = time.perf_counter()
    sim.advance(20); sim.sync(); t2 = time.perf_counter()
    sim.download_ptr(host.data_ptr(), P); t3 = time.perf_counter()
    print(f"frame {f}: upload {1e3*(t1-t0):.1f} ms, 20 substeps {1e3*(t2-t1):.1f} ms, download {1e3*(t3-t2):.1f} ms")
# raw copies
d = torch.empty(P * 104, dtype=torch.uint8, device="cuda")
for _ in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter(); d.copy_(host, non_blocking=True); torch.cuda.synchronize(); t1 = time.perf_counter()
    host.copy_(d, non_blocking=True); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"raw H2D {1e3*(t1-t0):.1f} ms ({P*104/(t1-t0)/1e9:.1f} GB/s), D2H {1e3*(t2-t1):.1f} ms ({P*104/(t2-t1)/1e9:.1f} GB/s)")
