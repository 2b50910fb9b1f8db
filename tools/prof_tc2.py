"""ncu target: the two-tile tensor-core solve (nfc_tc2.cuh) on 262144 columns of the config-2 distribution, 129 saves."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["NFC_TC_TILES"] = "2"
import nfc_b200
syn = nfc_b200.synthetic
NN = nfc_b200.named_chain("dense-default", seed=0)
for p in NN.params(): p *= 1e-5
theta, _ = nfc_b200.destructure(NN)
dev = torch.device("cuda:0")
B = int(os.environ.get("NFC_PROF_B", "262144"))
d = syn.make_columns("ocean", B=B, seed=1, K_CA=2.0)
T0, colp = torch.from_numpy(d["T0"]).to(dev), torch.from_numpy(d["colp"]).to(dev)
Tout = torch.empty((B, 129, 32), device=dev); st = torch.zeros(B, dtype=torch.int32, device=dev)
h = nfc_b200.Handle(NN.conv_width, NN.dense_spec, 1, syn.saveat(129)); h.set_weights(theta)
for i in range(3):
    h.solve_dev(T0, colp, d["glob"], Tout, None, None, st); c = h.counters()
print(c["last_kernel_ms"], c["steps_accepted"] + c["steps_rejected"], flush=True)
