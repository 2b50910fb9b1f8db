"""Per-column attempted-step distribution of the bench workload (config 2) -- what bounds the makespan of a persistent solve kernel."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nfc_b200
from oracle import nde_oracle as O
syn = nfc_b200.synthetic
cw, layers = O.chain_spec()
dev = torch.device("cuda", 0)
theta = O.glorot_theta(cw, layers, 0, 1e-5)
h = nfc_b200.Handle(cw, layers, 1, syn.saveat(129)); h.set_weights(theta)
B = 65536
d = syn.make_columns("ocean", B=B, seed=1)
T0 = torch.from_numpy(d["T0"]).to(dev); cp = torch.from_numpy(d["colp"]).to(dev)
out = torch.empty((B, 129, 32), device=dev)
na, nr, st = (torch.zeros(B, dtype=torch.int32, device=dev) for _ in range(3))
h.solve_dev(T0, cp, d["glob"], out, na, nr, st); torch.cuda.synchronize()
n = (na + nr).cpu().numpy()
print("steps/column: mean %.1f" % n.mean(), "percentiles 50/90/99/99.9/100:", [int(np.percentile(n, p)) for p in (50, 90, 99, 99.9, 100)], "min", n.min())
s = np.sort(n)[::-1]
for k in (1, 64, 640, 1280, 2560, 6553, 18944, 37888): print(f"  rank {k}: {s[k-1]} steps; head work share {s[:k].sum()/s.sum():.3f}")
