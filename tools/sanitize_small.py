"""Smallest run that touches every kernel (for compute-sanitizer memcheck)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nfc_b200
from oracle import nde_oracle as O
syn = nfc_b200.synthetic
d = syn.make_columns("ocean", B=200, seed=3)
sv = syn.saveat(5)
for arch, split in (("dense-default", "fp16"), ("dense-default", "bf16"), ("conv-2", "fp16"), ("dense-deeper", "fp16")):
    os.environ["NFC_TC_SPLIT"] = split
    cw, layers = O.chain_spec(arch)
    theta = O.glorot_theta(cw, layers, 0, 0.3)
    h = nfc_b200.Handle(cw, layers, 1, sv, adjoint_max_steps=256)
    h.set_weights(theta)
    f = h.rhs(d["T0"], d["colp"], d["glob"])
    r = h.solve(d["T0"], d["colp"], d["glob"])
    w = h.diagnose_flux(r["T"], d["colp"])
    g = h.loss_grad(d["T0"][:70], d["colp"][:70], d["glob"], r["T"][:70] * 0.99)      # <= 148 columns: one-column-per-CTA forward + backward kernels (default Chain)
    h.set_option("adj_small_max", 0)
    if arch == "dense-default":
        h.set_option("fwd_small_max", 0)
    g2 = h.loss_grad(d["T0"][:70], d["colp"][:70], d["glob"], r["T"][:70] * 0.99)    # batch kernels on the same columns
    Ts, wTs = syn.reference_scalings()
    e = h.embedded_solve(Ts.inv(d["T0"][:20]), wTs.inv(d["colp"][:20, 1]), [Ts.mu, Ts.sigma, wTs.mu, wTs.sigma], 256.0, 600.0, 1.0, 1e-3, 6, 2)
    print(arch, split, "ok", float(np.abs(f).max()), int(r["status"].max()), float(g["loss"]), e.shape, flush=True)
    h.close()
os.environ["NFC_TC"] = "0"
cw, layers = O.chain_spec()
h = nfc_b200.Handle(cw, layers, 1, sv); h.set_weights(O.glorot_theta(cw, layers, 0, 0.3))
r = h.solve(d["T0"], d["colp"], d["glob"]); print("simt ok", int(r["status"].max()), flush=True)
