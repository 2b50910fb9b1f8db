"""Adjoint first-contact: GPU loss_grad vs C oracle continuous adjoint, with per-layer breakdown."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nfc_b200
from oracle import nde_oracle as O, c_oracle as CO
syn = nfc_b200.synthetic
d = syn.make_columns("training9")
cw, layers = O.chain_spec()
def rel(a, b): return np.linalg.norm(a.astype(np.float64) - b) / np.linalg.norm(b)
for (scale, K, fixed) in [(1.0, 0.0, 1 / 256), (1.0, 2.0, 1 / 256), (1.0, 0.0, 0.0), (1.0, 2.0, 0.0), (0.3, 2.0, 0.0)]:
    theta = O.glorot_theta(cw, layers, 0, scale)
    sv = syn.saveat(33)
    colp = d["colp"].copy(); colp[:, 2] = K
    cp10 = colp.copy(); cp10[:, 2] = 10.0
    Ttrue = CO.solve(d["T0"], np.zeros_like(theta), cw, layers, cp10, d["glob"], sv, reltol=1e-6, abstol=1e-8, nthreads=8)["T"]
    h = nfc_b200.Handle(cw, layers, 1, sv, fixed_dt=fixed, adjoint_max_steps=512)
    h.set_weights(theta)
    t0 = time.time(); r = h.loss_grad(d["T0"], colp, d["glob"], Ttrue); t1 = time.time()
    o = CO.loss_grad_continuous(d["T0"], theta, cw, layers, colp, d["glob"], sv, Ttrue, fixed_dt=fixed, nthreads=8)
    print(f"scale {scale} K {K} fixed {fixed:.5f}: loss gpu {r['loss']:.8e} cpu {o['loss']:.8e} | grad rel {rel(r['grad'], o['grad'].astype(np.float64)):.3e} "
          f"|g| {np.linalg.norm(o['grad']):.3e} status {np.bincount(r['status'])} {t1-t0:.3f}s", flush=True)
    off = 0
    for (i, oo, a) in layers:
        for nm, n in (("W", i * oo), ("b", oo)):
            sl = slice(off, off + n); off += n
            print(f"    {nm}[{i}x{oo}] rel {rel(r['grad'][sl], o['grad'][sl].astype(np.float64)):.3e}  |g| {np.linalg.norm(o['grad'][sl]):.3e}", flush=True)
    print("    counters", {k: v for k, v in h.counters().items() if 'adj' in k or 'steps' in k}, "oracle adj steps", int(o['adj_naccept'].sum()), int(o['adj_nreject'].sum()), flush=True)
# throughput
B = 16384
d = syn.make_columns("ocean", B=B, seed=2)
sv = syn.saveat(129)
theta = O.glorot_theta(cw, layers, 0, 1e-5)
h = nfc_b200.Handle(cw, layers, 1, sv)
h.set_weights(theta)
Ttrue = np.repeat(d["T0"][:, None, :], 129, axis=1) + 0.01
for i in range(2):
    t0 = time.time(); r = h.loss_grad(d["T0"], d["colp"], d["glob"], Ttrue); t1 = time.time()
    c = h.counters()
    print(f"B={B}: wall {t1-t0:.3f}s fwd(record) {c['last_kernel_ms']:.1f} ms adjoint-total {c['last_adjoint_ms']:.1f} ms, fwd steps {c['steps_accepted']+c['steps_rejected']}, "
          f"adj steps {c['adj_steps_accepted']}+{c['adj_steps_rejected']}, loss {r['loss']:.6e}, status {np.bincount(r['status'])}", flush=True)
