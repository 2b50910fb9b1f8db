"""First-contact GPU sanity: RHS parity, small solve vs C oracle, timing of one 65,536-column solve."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nfc_b200
from oracle import nde_oracle as O, c_oracle as CO
syn = nfc_b200.synthetic
def rel(a, b):
    a = a.reshape(a.shape[0], -1).astype(np.float64); b = b.reshape(b.shape[0], -1).astype(np.float64)
    return np.linalg.norm(a - b, axis=1) / np.linalg.norm(b, axis=1)
cw, layers = O.chain_spec()
theta = O.glorot_theta(cw, layers, 0, 1.0)
d = syn.make_columns("ocean", B=1000, seed=3)
sv = syn.saveat(17)
h = nfc_b200.Handle(cw, layers, 1, sv)
h.set_weights(theta)
f = h.rhs(d["T0"], d["colp"], d["glob"])
ref = O.rhs(d["T0"], theta, cw, layers, d["colp"], d["glob"])
print("rhs rel err max", rel(f, ref).max(), flush=True)
t0 = time.time(); r = h.solve(d["T0"], d["colp"], d["glob"]); t1 = time.time()
print("solve 1000 cols: %.3fs" % (t1 - t0), "status", np.bincount(r["status"]), "acc", r["naccept"][:8], "rej", r["nreject"][:8], flush=True)
c = CO.solve(d["T0"], theta, cw, layers, d["colp"], d["glob"], sv, nthreads=8)
print("traj rel err vs C oracle max", rel(r["T"], c["T"]).max(), "steps gpu/cpu", (r["naccept"]+r["nreject"]).sum(), (c["naccept"]+c["nreject"]).sum(), flush=True)
print("counters", h.counters(), flush=True)
print("fp32 peak TF", nfc_b200.measure_fp32_peak(0, 50.0), flush=True)
import torch
B = 65536
d = syn.make_columns("ocean", B=B, seed=1)
sv = syn.saveat(129)
h2 = nfc_b200.Handle(cw, layers, 1, sv)
h2.set_weights(O.glorot_theta(cw, layers, 0, 1e-5))
dev = torch.device("cuda:0")
T0, colp = torch.from_numpy(d["T0"]).to(dev), torch.from_numpy(d["colp"]).to(dev)
Tout = torch.empty((B, 129, 32), device=dev)
st = torch.zeros(B, dtype=torch.int32, device=dev)
for i in range(3):
    h2.solve_dev(T0, colp, d["glob"], Tout, None, None, st)
    cnt = h2.counters()
    steps = cnt["steps_accepted"] + cnt["steps_rejected"]
    print(f"65536 cols: kernel {cnt['last_kernel_ms']:.2f} ms, steps {steps}, {steps/cnt['last_kernel_ms']*1e3:.3e} col-steps/s, "
          f"{steps*293376/cnt['last_kernel_ms']*1e3/1e12:.2f} TFLOP/s, bad status {int((st!=0).sum())}", flush=True)
