"""Per-evaluation cost of the tensor-core solve kernel: uniform workload (fixed dt, one tile per SM) and large batches."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nfc_b200
syn = nfc_b200.synthetic
NN = nfc_b200.named_chain("dense-default", seed=0)
for p in NN.params(): p *= 1e-5
theta, _ = nfc_b200.destructure(NN)
dev = torch.device("cuda:0")
def run(B, n_save, fixed_dt, tag):
    d = syn.make_columns("ocean", B=B, seed=1)
    h = nfc_b200.Handle(NN.conv_width, NN.dense_spec, 1, syn.saveat(n_save), fixed_dt=fixed_dt)
    h.set_weights(theta)
    T0, colp = torch.from_numpy(d["T0"]).to(dev), torch.from_numpy(d["colp"]).to(dev)
    Tout = torch.empty((B, n_save, 32), device=dev)
    st = torch.zeros(B, dtype=torch.int32, device=dev)
    for i in range(2):
        h.solve_dev(T0, colp, d["glob"], Tout, None, None, st)
        c = h.counters()
    steps = c["steps_accepted"] + c["steps_rejected"]
    ms = c["last_kernel_ms"]
    print(f"{tag}: B={B} n_save={n_save} fixed_dt={fixed_dt}: {ms:.2f} ms, steps {steps}, {steps/ms*1e3:.3e} col-steps/s", flush=True)
    return ms, steps
for tc in ("1", "0"):
    os.environ["NFC_TC"] = tc
    ms, steps = run(148 * 128, 2, 1 / 256, f"TC={tc} uniform")
    print(f"    -> {ms*1e-3*1.965e9/(256*6):.0f} cycles per Chain evaluation per tile of {128 if tc=='1' else 64} columns (at 1.965 GHz)", flush=True)
    run(148 * 128 * 4, 2, 1 / 256, f"TC={tc} uniform x4")
    run(262144, 129, 0.0, f"TC={tc} adaptive")
    run(1048576, 17, 0.0, f"TC={tc} adaptive")
