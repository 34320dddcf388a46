"""Summarises an .ncu-rep (first kernel) into the handful of metrics DESIGN.md/profiles cite.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [particles]"""
import csv, io, subprocess, sys

rep = sys.argv[1]
P = float(sys.argv[2]) if len(sys.argv) > 2 else 67108864.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
for vals in rows[2:]:
    m = dict(zip(hdr, vals))
    u = dict(zip(hdr, units))
    def g(k, d=0.0):
        try: return float(m[k].replace(",", ""))
        except Exception: return d
    def scale(k):  # normalise to base units
        v = g(k); un = u.get(k, "")
        return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1, "second":1, "msecond":1e-3, "usecond":1e-6}.get(un, 1)
    t = scale("gpu__time_duration.sum")
    rd, wr = scale("dram__bytes_read.sum"), scale("dram__bytes_write.sum")
    inst = g("smsp__inst_executed.sum")
    print(f"kernel: {m.get('Kernel Name','?')[:90]}")
    print(f"  duration {t*1e3:.3f} ms | regs {g('launch__registers_per_thread'):.0f} | achieved occupancy {g('sm__warps_active.avg.pct_of_peak_sustained_active'):.1f}% | SM clock {g('sm__cycles_elapsed.avg.per_second'):.3f} GHz")
    print(f"  dram read {rd/1e9:.3f} GB write {wr/1e9:.3f} GB -> {(rd+wr)/t/1e9:.0f} GB/s ({g('dram__throughput.avg.pct_of_peak_sustained_elapsed'):.1f}% of ncu peak)")
    print(f"  warp-instr {inst:.4g} -> {inst*32/P:.0f} thread-instr/particle | issue active {g('smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f}% | sm throughput {g('sm__throughput.avg.pct_of_peak_sustained_elapsed'):.1f}%")
    print(f"  L1 hit {g('l1tex__t_sector_hit_rate.pct'):.1f}% | L2 hit {g('lts__t_sector_hit_rate.pct'):.1f}% | global RED instr {g('smsp__inst_executed_op_global_red.sum'):.4g} | lsu pipe {g('sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active'):.1f}%")
    stalls = sorted(((g(k), k) for k in hdr if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio")), reverse=True)[:6]
    for v, k in stalls:
        print(f"  stall {k.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio','')}: {v:.2f}")
    extra = ["l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active","l1tex__throughput.avg.pct_of_peak_sustained_elapsed","lts__throughput.avg.pct_of_peak_sustained_elapsed"]
    for k in extra:
        if k in m: print(f"  {k}: {m[k]} {u.get(k,'')}")
