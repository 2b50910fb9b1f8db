"""Two-tile tensor-core solve (nfc_tc2.cuh) against the one-tile kernel: bit-for-bit agreement and kernel time."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nfc_b200
from oracle import nde_oracle as O
syn = nfc_b200.synthetic
cw, layers = O.chain_spec()
sv = syn.saveat(129)
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
def handle(tiles, theta):
    os.environ["NFC_TC_TILES"] = str(tiles)
    h = nfc_b200.Handle(cw, layers, 1, sv); h.set_weights(theta)
    return h
def run(h, d, reps=3):
    B = d["T0"].shape[0]
    T0 = torch.from_numpy(d["T0"]).to(dev); cp = torch.from_numpy(d["colp"]).to(dev)
    out = torch.full((B, 129, 32), float("nan"), device=dev) if B <= 300000 else torch.empty((B, 129, 32), device=dev)
    na, nr, st = (torch.zeros(B, dtype=torch.int32, device=dev) for _ in range(3))
    ms = []
    for _ in range(reps):
        h.solve_dev(T0, cp, d["glob"], out, na, nr, st, stream=stream.cuda_stream)
        torch.cuda.synchronize()
        ms.append(h.counters()["last_kernel_ms"])
    c = h.counters()
    res = out.cpu().numpy(); del out
    return res, na.cpu().numpy(), nr.cpu().numpy(), st.cpu().numpy(), ms, c["steps_accepted"] + c["steps_rejected"]
for scale in [float(x) for x in os.environ.get("NFC_TC2_SCALES", "1e-5,0.3").split(",")]:
    theta = O.glorot_theta(cw, layers, 0, scale)
    h1, h2 = handle(1, theta), handle(2, theta)
    for B in [int(x) for x in os.environ.get("NFC_TC2_B", "1,129,300,4096,65536").split(",")]:
        d = syn.make_columns("ocean", B=B, seed=1)
        a = run(h1, d); b = run(h2, d)
        same = np.array_equal(a[0].view(np.uint32), b[0].view(np.uint32))
        rel = np.nanmax(np.abs(a[0] - b[0])) / np.nanmax(np.abs(a[0]))
        print(f"weights x{scale:g} B={B}: bitwise {same} (max rel diff {rel:.2e}, nan {int(np.isnan(b[0]).sum())}), nacc equal {np.array_equal(a[1], b[1])}, nrej equal {np.array_equal(a[2], b[2])}, "
              f"status equal {np.array_equal(a[3], b[3])} bad {int((b[3] != 0).sum())} | steps {a[5]} vs {b[5]} | one tile {a[4][-1]:.3f} ms, two tiles {b[4][-1]:.3f} ms "
              f"({b[5] / b[4][-1] / 1e3:.3e} col-steps/s)", flush=True)
