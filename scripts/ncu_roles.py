"""Per-role accounting of an .ncu-rep of agg_tc_kernel: SASS lines are bucketed by executed count
(producer / weight warps execute each line once per tile-warp), with stall reasons and spin loops listed."""
import csv, io, re, subprocess, sys
from collections import Counter
rep = sys.argv[1]
tiles = int(sys.argv[2]) if len(sys.argv) > 2 else 50000
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0          # kernel index inside the report
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
rows = rows[starts[which]:starts[which + 1]]
hdr = rows[1]; data = [r for r in rows[2:] if len(r) == len(hdr)]
ie = hdr.index("Instructions Executed"); sm = hdr.index("# Samples"); so = hdr.index("Source")
stall = [i for i, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
print("kernel:", rows[0][1][:60])
tot_s = sum(int(r[sm]) for r in data if r[sm].isdigit())
segs = []; cur = None
for idx, r in enumerate(data):
    try: e = int(r[ie])
    except ValueError: continue
    if 7.9 * tiles <= e <= 8.1 * tiles:
        if cur and idx - cur[1] <= 40: cur[1] = idx
        else: cur = [idx, idx]; segs.append(cur)
print(f"total samples {tot_s} (per warp {tot_s/20:.0f})")
for a, b in segs:
    n = 0; smp = 0; ops = Counter(); st = Counter()
    for r in data[a:b + 1]:
        try: e = int(r[ie])
        except ValueError: continue
        smp += int(r[sm])
        for i in stall: st[hdr[i]] += int(r[i] or 0)
        if 7.9 * tiles <= e <= 8.1 * tiles:
            n += 1; ops[re.sub(r"^@!?U?P\d+\s+", "", r[so].strip()).split()[0]] += 1
    print(f"seg {a}-{b}: {n} lines x8 warps/tile, samples {smp} ({100*smp/tot_s:.1f}%)", ops.most_common(8), st.most_common(4))
print("-- hot lines")
hot = sorted(((int(r[sm]), i, r) for i, r in enumerate(data) if r[sm].isdigit()), reverse=True)[:25]
for s, i, r in hot:
    top = sorted(((int(r[k] or 0), hdr[k]) for k in stall), reverse=True)[:2]
    print(f"  {i:5d} {100*s/tot_s:5.1f}% exec {r[ie]:>9s} {r[so].strip()[:70]:70s} {top}")
