"""Does this box overlap a host->device and a device->host copy?  (explains bench.py's e2e ceiling)"""
import time, torch
n = 7 * 1024**3 // 4
a = torch.empty(n, dtype=torch.float32, pin_memory=True); a.zero_()
b = torch.empty(n, dtype=torch.float32, pin_memory=True); b.zero_()
da = torch.empty(n, dtype=torch.float32, device="cuda")
db = torch.zeros(n, dtype=torch.float32, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h):
    torch.cuda.synchronize(); t = time.perf_counter()
    if h2d:
        with torch.cuda.stream(s1): da.copy_(a, non_blocking=True)
    if d2h:
        with torch.cuda.stream(s2): b.copy_(db, non_blocking=True)
    torch.cuda.synchronize(); return time.perf_counter() - t
for _ in range(2):
    t1, t2, t3 = run(True, False), run(False, True), run(True, True)
    gb = n * 4 / 1e9
    print(f"H2D alone {t1*1e3:.1f} ms ({gb/t1:.1f} GB/s) | D2H alone {t2*1e3:.1f} ms ({gb/t2:.1f} GB/s) | both {t3*1e3:.1f} ms ({2*gb/t3:.1f} GB/s combined)")
