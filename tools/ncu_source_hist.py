"""Dev helper: per-opcode executed-instruction + stall histogram from `ncu --page source --csv` output.
usage: ncu -i X.ncu-rep --page source --csv > src.csv ; python tools/ncu_source_hist.py src.csv [top_lines]"""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ci = {h: i for i, h in enumerate(hdr)}
tot = 0
byop, stall = collections.Counter(), collections.Counter()
stall_kinds = collections.Counter()
data = []
kinds = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    n = int(r[ci["Instructions Executed"]])
    m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_]+)(\.[A-Z0-9_.]*)?", r[ci["Source"]])
    op = m.group(1) if m else "?"
    if op in ("LDS", "STS", "LDG", "STG") and m.group(2) and "128" in m.group(2):
        op += ".128"
    byop[op] += n
    tot += n
    s = int(r[ci["Warp Stall Sampling (All Samples)"]])
    stall[op] += s
    for k in kinds:
        stall_kinds[k] += int(r[ci[k]])
    data.append((n, s, r[ci["Source"]].strip(), r[ci["L1 Wavefronts Shared"]], r[ci["L1 Wavefronts Shared Ideal"]]))
print("total warp-instructions", tot, " stall samples", sum(stall.values()))
for op, n in byop.most_common(24):
    print(f"{op:10s} {n:>14d} {100*n/tot:5.1f}%   stall samples {stall[op]:>8d} {100*stall[op]/max(1,sum(stall.values())):5.1f}%")
print({k: v for k, v in stall_kinds.most_common(8)})
if len(sys.argv) > 2:
    print("--- top lines by stall samples")
    for n, s, src, wf, wfi in sorted(data, key=lambda x: -x[1])[: int(sys.argv[2])]:
        print(f"{s:8d} {n:>12d} wf {wf:>12s}/{wfi:>12s}  {src}")
