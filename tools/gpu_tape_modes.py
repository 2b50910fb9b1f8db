"""Adjoint tape layouts (option tape_mode): loss + gradient time and tape memory, dense against compact, at 65 536 columns (one dense
chunk) and at config 4's 1 048 576 columns on ONE GPU (dense: 8 chunks of 48 GiB; compact: 2 chunks at 48 GiB, 1 at 100 GiB)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nfc_b200
dev = torch.device("cuda", 0)
syn = nfc_b200.synthetic
NN = nfc_b200.named_chain("dense-default", seed=0)
for p in NN.params(): p *= 1e-5
theta, _ = nfc_b200.destructure(NN)
sv = syn.saveat(129)
s = torch.cuda.Stream(device=dev); torch.cuda.set_stream(s)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for B, cases in ((65536, ((1, 48), (2, 48))), (int(os.environ.get("NFC_TAPE_BIG", "1048576")), ((1, 48), (2, 48), (2, 100)))):
    d = syn.make_columns("ocean", B=B, seed=2, K_CA=2.0)
    T0 = torch.from_numpy(d["T0"]).to(dev); colp = torch.from_numpy(d["colp"]).to(dev)
    ref = None
    for mode, gib in cases:
        h = nfc_b200.Handle(NN.conv_width, NN.dense_spec, 1, sv)
        h.set_weights(theta)
        h.set_option("tape_mode", mode); h.set_option("tape_gib", gib)
        Tout = torch.empty((B, 129, 32), device=dev); st = torch.zeros(B, dtype=torch.int32, device=dev)
        h.solve_dev(T0, colp, d["glob"], Tout, None, None, st, stream=s.cuda_stream)
        Ttrue = Tout.mul_(0.99).add_(0.01); del Tout
        g = torch.empty(h.n_params, device=dev); ls = torch.zeros(2, dtype=torch.float64, device=dev)
        free0 = torch.cuda.mem_get_info()[0]
        ts = []
        for it in range(4):
            torch.cuda.synchronize(); e0.record()
            h.loss_grad_dev(T0, colp, d["glob"], Ttrue, ls, g, st, stream=s.cuda_stream)
            e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        free1 = torch.cuda.mem_get_info()[0]
        c = h.counters()
        gn = g.double()
        if ref is None: ref = (float(ls[0]), gn.clone())
        print(f"{B} columns, tape_mode {mode} ({'dense' if mode == 1 else 'compact'}), budget {gib} GiB: loss+grad {ts[0]:.1f} / {ts[1]:.1f} / {ts[2]:.1f} / {ts[3]:.1f} ms, "
              f"scratch {(free0 - free1) / 2**30:.1f} GiB, launches {c['kernel_launches']}, bad status {int((st != 0).sum())}, "
              f"loss sum {float(ls[0]):.9e} (rel diff to first {abs(float(ls[0]) - ref[0]) / ref[0]:.1e}), grad rel diff {float((gn - ref[1]).norm() / ref[1].norm()):.1e}", flush=True)
        del h, Ttrue, g, ls, st
        torch.cuda.empty_cache()
