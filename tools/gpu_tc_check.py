"""tensor-core (tcgen05) RHS vs float64 oracle and vs the SIMT kernel."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nfc_b200
from oracle import nde_oracle as O
syn = nfc_b200.synthetic
cw, layers = O.chain_spec()
def rel(a, b):
    a = a.reshape(a.shape[0], -1).astype(np.float64); b = b.reshape(b.shape[0], -1).astype(np.float64)
    return np.linalg.norm(a - b, axis=1) / np.linalg.norm(b, axis=1)
for B in (128, 1, 300, 40000):
    for scale in (1.0, 1e-5):
        d = syn.make_columns("ocean", B=B, seed=3)
        rng = np.random.default_rng(B)
        T = (d["T0"] + 0.3 * rng.standard_normal(d["T0"].shape)).astype(np.float32)
        theta = O.glorot_theta(cw, layers, 0, scale); theta[-31:] = 0.1 * scale
        ref = O.rhs(T, theta, cw, layers, d["colp"], d["glob"])
        os.environ["NFC_TC"] = "1"
        h = nfc_b200.Handle(cw, layers, 1, syn.saveat(9)); h.set_weights(theta)
        t0 = time.time(); f = h.rhs(T, d["colp"], d["glob"]); t1 = time.time()
        os.environ["NFC_TC"] = "0"
        h0 = nfc_b200.Handle(cw, layers, 1, syn.saveat(9)); h0.set_weights(theta)
        f0 = h0.rhs(T, d["colp"], d["glob"])
        e = rel(f, ref)
        print(f"B={B} scale={scale}: TC rel err max {e.max():.3e} mean {e.mean():.3e} | SIMT rel err max {rel(f0, ref).max():.3e} | nan {np.isnan(f).sum()} | {t1-t0:.4f}s", flush=True)
        if e.max() > 1e-3:
            i = int(np.argmax(e)); print("   worst col", i, "tc", f[i, :6], "ref", ref[i, :6], flush=True)
