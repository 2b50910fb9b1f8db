"""End-to-end nfc_solve (pinned host buffers in and out) of the bench workload against the number of pipeline pieces."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nfc_b200
from oracle import nde_oracle as O
syn = nfc_b200.synthetic
cw, layers = O.chain_spec()
theta = O.glorot_theta(cw, layers, 0, 1e-5)
B = 65536
d = syn.make_columns("ocean", B=B, seed=1, K_CA=2.0)
hT0 = torch.from_numpy(d["T0"]).pin_memory(); hcolp = torch.from_numpy(d["colp"]).pin_memory()
hout = torch.empty((B, 129, 32), dtype=torch.float32).pin_memory()
for pieces in os.environ.get("NFC_PIECES_LIST", "1,2,3,4,6,8").split(","):
    os.environ["NFC_E2E_PIECES"] = pieces
    h = nfc_b200.Handle(cw, layers, 1, syn.saveat(129)); h.set_weights(theta)
    for _ in range(3):
        r = h.solve(hT0.numpy(), hcolp.numpy(), d["glob"], out=hout.numpy())
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(4):
        r = h.solve(hT0.numpy(), hcolp.numpy(), d["glob"], out=hout.numpy())
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 4
    steps = int((r["naccept"] + r["nreject"]).sum())
    print(f"pieces {pieces}: {dt*1e3:.2f} ms per call, {steps/dt:.3e} column-steps/s end to end, D2H {hout.numel()*4/dt/1e9:.1f} GB/s equivalent", flush=True)
    del h
