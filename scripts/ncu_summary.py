"""Summarise an .ncu-rep: key raw metrics + top stall lines from the source page.  usage: ncu_summary.py rep [kernel-regex]"""
import csv, io, subprocess, sys, re
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__registers_per_thread", "smsp__inst_executed.sum",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.avg", "smsp__cycles_active.avg", "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_xu.sum"]
for r in rows[2:]:
    print("==", r[hdr.index("Kernel Name")])
    for k in keys:
        if k in hdr:
            i = hdr.index(k); print(f"   {k:80s} {r[i]:>16s} {units[i]}")
    for i, h in enumerate(hdr):
        if "smsp__average_warp" in h and "issue_stalled" in h and "not_issued" not in h and h.endswith(".ratio"):
            try:
                if float(r[i]) > 0.3: print(f"   {h:80s} {r[i]:>16s}")
            except ValueError: pass
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
if len(rows) > 2:
    # find header
    hi = next(i for i, r in enumerate(rows) if "Source" in r and any("Sampling" in c for c in r))
    hdr = rows[hi]
    si = hdr.index("Source"); smp = next(i for i, c in enumerate(hdr) if c.startswith("# Samples") or c == "Warp Stall Sampling (All Samples)" or "Sampling (All" in c)
    data = []
    for r in rows[hi + 1:]:
        try: data.append((int(r[smp]), r[si][:150], r))
        except (ValueError, IndexError): pass
    tot = sum(d[0] for d in data) or 1
    print(f"-- top source lines by warp-stall samples (total {tot})")
    stall_cols = [i for i, c in enumerate(hdr) if c.startswith("stall_")]
    for n, s, r in sorted(data, key=lambda d: -d[0])[:28]:
        top = sorted(((int(r[i]) if r[i].isdigit() else 0, hdr[i]) for i in stall_cols), reverse=True)[:2]
        print(f"   {100*n/tot:5.1f}%  {s.strip()[:110]:110s} {top}")
