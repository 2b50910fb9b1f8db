import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nfc_b200
syn = nfc_b200.synthetic
NN = nfc_b200.named_chain("dense-default", seed=0)
for p in NN.params(): p *= 1e-5
theta, _ = nfc_b200.destructure(NN)
B = 148 * 64
d = syn.make_columns("ocean", B=B, seed=2)
sv = syn.saveat(17)
h = nfc_b200.Handle(NN.conv_width, NN.dense_spec, 1, sv, fixed_dt=1 / 128)
h.set_weights(theta)
Ttrue = np.repeat(d["T0"][:, None, :], 17, axis=1) + 0.01
for i in range(2):
    r = h.loss_grad(d["T0"], d["colp"], d["glob"], Ttrue); c = h.counters()
print(c["last_adjoint_ms"], c["adj_steps_accepted"], flush=True)
