"""Per-role accounting of an .ncu-rep of agg_tc_kernel / agg_bf16_kernel: the SASS is cut at the USETMAXREG instructions
(producers | issuers | weights) and samples / stall reasons / executed instructions are summed per region.
    python scripts/ncu_regions.py report.ncu-rep [kernel index] [top lines per region]"""
import csv, io, subprocess, sys
from collections import Counter
rep = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 12
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
rows = rows[starts[which]:starts[which + 1]]
hdr = rows[1]; data = [r for r in rows[2:] if len(r) == len(hdr)]
ie = hdr.index("Instructions Executed"); sm = hdr.index("# Samples"); so = hdr.index("Source")
stall = [i for i, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
print("kernel:", rows[0][1][:70])
cuts = [i for i, r in enumerate(data) if "USETMAXREG" in r[so]]
bounds = [0] + cuts + [len(data)]
tot = sum(int(r[sm] or 0) for r in data); toti = sum(int(r[ie] or 0) for r in data)
print(f"total samples {tot}, warp instructions {toti}")
for a, b in zip(bounds[:-1], bounds[1:]):
    seg = data[a:b]
    s = sum(int(r[sm] or 0) for r in seg); n = sum(int(r[ie] or 0) for r in seg)
    st = Counter()
    for r in seg:
        for i in stall: st[hdr[i]] += int(r[i] or 0)
    print(f"\n== lines {a}-{b} ({data[a][so].strip()[:40]}): samples {s} ({100*s/max(tot,1):.1f}%), instructions {n} ({100*n/max(toti,1):.1f}%)")
    print("   stalls:", [(k, v) for k, v in st.most_common(6)])
    hot = sorted(((int(r[sm] or 0), i, r) for i, r in enumerate(seg)), reverse=True)[:top]
    for sm_, i, r in hot:
        tp = sorted(((int(r[k] or 0), hdr[k]) for k in stall), reverse=True)[:2]
        print(f"   {a+i:5d} {100*sm_/max(tot,1):5.1f}% exec {r[ie]:>9s} {r[so].strip()[:64]:64s} {tp}")
