"""Gradient of the 9-simulation problem: GPU kernels and the C twin against the converged float64 gradient (autograd on a fine fixed grid)."""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nfc_b200
from oracle import c_oracle as CO, nde_oracle as O
syn = nfc_b200.synthetic
def rel(a, b): return float(np.linalg.norm(np.float64(a) - np.float64(b)) / np.linalg.norm(np.float64(b)))
d = syn.make_columns("training9"); cw, layers = O.chain_spec("dense-default")
sv = syn.saveat(33)
for scale, K in ((0.3, 2.0), (1.0, 0.0), (1.0, 2.0)):
    theta = O.glorot_theta(cw, layers, 0, scale)
    colp = d["colp"].copy(); colp[:, 2] = K
    cp10 = colp.copy(); cp10[:, 2] = 10.0
    Ttrue = CO.solve(d["T0"], np.zeros_like(theta), cw, layers, cp10, d["glob"], sv, reltol=1e-6, abstol=1e-8, nthreads=8)["T"]
    L, g, _ = O.loss_and_grad_autograd(d["T0"], theta, cw, layers, colp, d["glob"], sv, Ttrue, fixed_dt=1 / 1024)
    o = CO.loss_grad_continuous(d["T0"], theta, cw, layers, colp, d["glob"], sv, Ttrue, nthreads=8)
    print(f"scale {scale} K {K}: C twin vs truth {rel(o['grad'], g):.2e}", flush=True)
    for rt in (1e-5, 1e-6):
        h = nfc_b200.Handle(cw, layers, 1, sv, reltol=rt, abstol=rt * 1e-2, adjoint_max_steps=4096); h.set_weights(theta)
        r = h.loss_grad(d["T0"], colp, d["glob"], Ttrue)
        oc = CO.loss_grad_continuous(d["T0"], theta, cw, layers, colp, d["glob"], sv, Ttrue, reltol=rt, abstol=rt * 1e-2, nthreads=8)
        print(f"   reltol {rt:g}: GPU vs truth {rel(r['grad'], g):.2e}  C twin vs truth {rel(oc['grad'], g):.2e}  GPU vs C twin {rel(r['grad'], oc['grad']):.2e}", flush=True)
    for tag, opts in (("default", {}), ("record simt", {"tc_record": 0}), ("all simt", {"tc": 0}), ("small off", {"adj_small_max": 0}), ("small off, adj simt", {"adj_small_max": 0, "adj_tc": 0})):
        h = nfc_b200.Handle(cw, layers, 1, sv); h.set_weights(theta)
        for k, v in opts.items(): h.set_option(k, v)
        r = h.loss_grad(d["T0"], colp, d["glob"], Ttrue)
        c = h.counters()
        print(f"   GPU {tag:20s}: vs truth {rel(r['grad'], g):.2e}  vs C twin {rel(r['grad'], o['grad']):.2e}  loss rel {abs(r['loss'] - L) / L:.1e}  fwd steps {c['steps_accepted']}+{c['steps_rejected']} adj {c['adj_steps_accepted']}+{c['adj_steps_rejected']}", flush=True)
