"""Small-batch loss + gradient: the one-column-per-CTA backward kernel (nfc_adjoint_small.cuh) against the batch kernel."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nfc_b200
syn = nfc_b200.synthetic
NN = nfc_b200.named_chain("dense-default", seed=0)
for p in NN.params(): p *= 1e-5
theta, _ = nfc_b200.destructure(NN)
sv = syn.saveat(129)
def run(B, small_max):
    os.environ["NFC_ADJ_SMALL_MAX"] = str(small_max)
    d = syn.make_columns("training9") if B == 9 else syn.make_columns("ocean", B=B, seed=2)
    h = nfc_b200.Handle(NN.conv_width, NN.dense_spec, 1, sv); h.set_weights(theta)
    Ttrue = np.repeat(d["T0"][:, None, :], 129, axis=1) * np.float32(0.98)
    for i in range(3):
        t0 = time.perf_counter(); r = h.loss_grad(d["T0"], d["colp"], d["glob"], Ttrue); dt = time.perf_counter() - t0
    c = h.counters()
    return r, c, dt
for B in [int(x) for x in os.environ.get("NFC_ADJS_B", "9,64,296,1184,2368").split(",")]:
    a, ca, ta = run(B, 0)
    b, cb, tb = run(B, 1 << 30)
    rel = np.linalg.norm(a["grad"] - b["grad"]) / np.linalg.norm(a["grad"])
    print(f"B={B}: batch kernel {ca['last_adjoint_ms']:.2f} ms (call {ta*1e3:.1f}), one-column kernel {cb['last_adjoint_ms']:.2f} ms (call {tb*1e3:.1f}); "
          f"|g| {np.linalg.norm(a['grad']):.4e}, rel diff of gradients {rel:.2e}, loss {a['loss']:.6e} / {b['loss']:.6e}, "
          f"adj steps {ca['adj_steps_accepted']}+{ca['adj_steps_rejected']} / {cb['adj_steps_accepted']}+{cb['adj_steps_rejected']}, "
          f"bad {int((a['status']!=0).sum())}/{int((b['status']!=0).sum())}", flush=True)
