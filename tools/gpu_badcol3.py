import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nfc_b200
syn = nfc_b200.synthetic
d = syn.make_columns("ocean", B=1048576, seed=2, K_CA=2.0)
NN = nfc_b200.named_chain("dense-default", seed=0)
for p in NN.params(): p *= 1e-5
theta, _ = nfc_b200.destructure(NN)
sv = syn.saveat(129)
idx = [426987]
T0 = d["T0"][idx]; colp = d["colp"][idx]
Tt = np.repeat(T0[:, None, :], 129, axis=1) * np.float32(0.98)
h = nfc_b200.Handle(NN.conv_width, NN.dense_spec, 1, sv, max_iters=2000); h.set_weights(theta)
r = h.loss_grad(T0, colp, d["glob"], Tt)
print(r["status"], h.adjoint_steps(1))
tp = h.tape(0)
print("tape t:", tp[:, 160].tolist())
print("tape h:", tp[:, 161].tolist())
