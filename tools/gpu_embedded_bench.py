"""SURVEY row f1 measurement: embedded fixed-step path (explicit NN flux step + implicit convective adjustment, one Chain evaluation and one
32-unknown tridiagonal solve per column and step; Delta t = 600 s, 8 days = 1152 steps as in the LES) through nfc_embedded_solve."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nfc_b200
from oracle import nde_oracle as O
syn = nfc_b200.synthetic
cw, layers = O.chain_spec()
theta = O.glorot_theta(cw, layers, 0, 1e-5)
Ts, wTs = syn.reference_scalings()
for B in (4096, 65536, 262144):
    d = syn.make_columns("ocean", B=B, seed=3)
    T0 = Ts.inv(d["T0"]).astype(np.float32); Q = wTs.inv(d["colp"][:, 1]).astype(np.float32)
    h = nfc_b200.Handle(cw, layers, 1, syn.saveat(2)); h.set_weights(theta)
    n_steps = 1152
    for it in range(3):
        t0 = time.perf_counter()
        out = h.embedded_solve(T0, Q, [Ts.mu, Ts.sigma, wTs.mu, wTs.sigma], 256.0, 600.0, 10.0, 1e-3, n_steps, save_every=n_steps)
        dt = time.perf_counter() - t0
    rate = B * n_steps / dt
    print(f"B={B}: {dt*1e3:.1f} ms per call ({n_steps} steps, host buffers in / first+last profile out), {rate:.3e} column-steps/s, "
          f"{rate * 48896 / 1e12:.1f} TFLOP/s (one Chain evaluation = 48,896 FLOP per step), finite {bool(np.isfinite(out).all())}", flush=True)
