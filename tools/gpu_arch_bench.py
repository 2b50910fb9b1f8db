"""Forward solve of 65 536 columns for the five Chains: tensor-core kernel (where the Chain has one) against the FP32 SIMT kernel."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nfc_b200
syn = nfc_b200.synthetic
B = int(os.environ.get("NFC_PROF_B", "65536"))
d = syn.make_columns("ocean", B=B, seed=1)
dev = torch.device("cuda:0")
T0, colp = torch.from_numpy(d["T0"]).to(dev), torch.from_numpy(d["colp"]).to(dev)
Tout = torch.empty((B, 129, 32), device=dev)
st = torch.zeros(B, dtype=torch.int32, device=dev)
for arch in ("dense-default", "conv-2", "conv-4", "dense-deeper", "dense-wider"):
    NN = nfc_b200.named_chain(arch, seed=0)
    for p in NN.params(): p *= 1e-5
    theta, _ = nfc_b200.destructure(NN)
    h = nfc_b200.Handle(NN.conv_width, NN.dense_spec, 1, syn.saveat(129))
    h.set_weights(theta)
    out = []
    for tc in (1, 0):
        if tc == 1 and arch == "dense-wider":
            out.append("   (none)"); continue
        h.set_option("tc", tc)
        for i in range(2):
            h.solve_dev(T0, colp, d["glob"], Tout, None, None, st)
            c = h.counters()
        steps = c["steps_accepted"] + c["steps_rejected"]
        out.append(f"{c['last_kernel_ms']:8.2f} ms {steps / c['last_kernel_ms'] / 1e3:8.1f} M column-steps/s")
    print(f"{arch:14s} tensor cores {out[0]} | FP32 SIMT {out[1]}  (bad status {int((st != 0).sum())})", flush=True)
