"""SURVEY row f2 measurement: RKC2 (stabilised explicit, ROCK4's role) vs Tsit5 on the K_CA sweep of figureC (10.^range(-3,1,...)),
4096 columns per K_CA value, 129 saves, FP32 SIMT kernels for both and the tensor-core Tsit5 path for reference."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nfc_b200
from oracle import nde_oracle as O
syn = nfc_b200.synthetic
cw, layers = O.chain_spec()
sv = syn.saveat(129)
theta = O.glorot_theta(cw, layers, 0, 0.3)
B = int(os.environ.get('NFC_RKC_B', '16384'))
d = syn.make_columns("ocean", B=B, seed=6)
import torch
dev = torch.device("cuda", 0)
T0d = torch.from_numpy(d["T0"]).to(dev)
out = torch.empty((B, 129, 32), device=dev)
na, nr, st = (torch.zeros(B, dtype=torch.int32, device=dev) for _ in range(3))
stream = torch.cuda.Stream(device=dev)
def run(h, colp):
    cd = torch.from_numpy(colp).to(dev)
    for _ in range(2):
        h.solve_dev(T0d, cd, d["glob"], out, na, nr, st, stream=stream.cuda_stream)
        torch.cuda.synchronize()
    c = h.counters()
    return dict(T=out.cpu().numpy(), status=st.cpu().numpy()), c
os.environ["NFC_TC"] = "0"
hr = nfc_b200.Handle(cw, layers, 1, sv, alg="RKC2"); hr.set_weights(theta)
ht = nfc_b200.Handle(cw, layers, 1, sv); ht.set_weights(theta)
os.environ["NFC_TC"] = "1"
hc = nfc_b200.Handle(cw, layers, 1, sv); hc.set_weights(theta)
print("K_CA | RKC2: ms, evals/col, steps/col (rej%) | Tsit5 FP32: ms, evals/col | Tsit5 tensor-core: ms | max rel diff RKC2 vs Tsit5")
for K in [float(k) for k in os.environ.get('NFC_RKC_K', '0.001,0.1,1,2,5,10').split(',')]:
    colp = d["colp"].copy(); colp[:, 2] = K
    r, cr = run(hr, colp); t, ct = run(ht, colp); tc_, cc = run(hc, colp)
    diff = (np.linalg.norm((r["T"] - t["T"]).reshape(B, -1), axis=1) / np.linalg.norm(t["T"].reshape(B, -1), axis=1)).max()
    print(f"{K:6.3f} | [{cr['last_kernel_ms']*1e6/cr['rhs_evals']:.2f} vs {ct['last_kernel_ms']*1e6/ct['rhs_evals']:.2f} vs {cc['last_kernel_ms']*1e6/cc['rhs_evals']:.2f} ns/eval] | {cr['last_kernel_ms']:8.2f} {cr['rhs_evals']/B:7.0f} {(cr['steps_accepted']+cr['steps_rejected'])/B:6.0f} ({100*cr['steps_rejected']/(cr['steps_accepted']+cr['steps_rejected']):.0f}%) | "
          f"{ct['last_kernel_ms']:8.2f} {ct['rhs_evals']/B:7.0f} | {cc['last_kernel_ms']:8.2f} | {diff:.2e}  bad {int((r['status']!=0).sum())}/{int((t['status']!=0).sum())}")
