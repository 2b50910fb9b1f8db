"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time and share per kernel name."""
import collections, csv, sys
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]; ci = {h: i for i, h in enumerate(hdr)}
tot = collections.Counter(); cnt = collections.Counter()
for r in rows[1:]:
    if len(r) < len(hdr) or r[ci["Metric Name"]] != "gpu__time_duration.sum":
        continue
    v = float(r[ci["Metric Value"]].replace(",", ""))
    u = r[ci["Metric Unit"]]
    ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6)
    k = r[ci["Kernel Name"]][:110]
    tot[k] += ms; cnt[k] += 1
T = sum(tot.values())
print(f"# total {T:.2f} ms over {sum(cnt.values())} launches")
for k, v in tot.most_common():
    print(f"{v:10.3f} ms  {100*v/T:5.1f}%  x{cnt[k]:3d}  {k}")
