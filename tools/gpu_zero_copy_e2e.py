"""Experiment: nfc_solve end to end (65 536 columns x 129 saves, 1.08 GB of trajectories) with the solve kernel writing the trajectories
straight into PINNED HOST memory (zero copy: streaming stores over PCIe while the solve runs) against the library's staged path
(device buffer + pipelined cudaMemcpyAsync)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nfc_b200
dev = torch.device("cuda", 0)
syn = nfc_b200.synthetic
B = int(os.environ.get("NFC_B", "65536"))
d = syn.make_columns("ocean", B=B, seed=1, K_CA=2.0)
NN = nfc_b200.named_chain("dense-default", seed=0)
for p in NN.params(): p *= 1e-5
theta, _ = nfc_b200.destructure(NN)
sv = syn.saveat(129)
h = nfc_b200.Handle(NN.conv_width, NN.dense_spec, 1, sv); h.set_weights(theta)
hT0 = torch.from_numpy(d["T0"]).pin_memory(); hcolp = torch.from_numpy(d["colp"]).pin_memory()
hout = torch.empty((B, 129, 32), dtype=torch.float32).pin_memory()
s = torch.cuda.Stream(device=dev); torch.cuda.set_stream(s)
def staged():
    return h.solve(hT0.numpy(), hcolp.numpy(), d["glob"], out=hout.numpy())
dT0 = torch.empty((B, 32), device=dev); dcolp = torch.empty((B, 3), device=dev)
na = torch.zeros(B, dtype=torch.int32, device=dev); nr = torch.zeros_like(na); st = torch.zeros_like(na)
class HostOut:          # solve_dev only needs data_ptr(): the pinned buffer's address is valid on the device (unified addressing)
    def __init__(self, t): self.t = t
    def data_ptr(self): return self.t.data_ptr()
def zero_copy():
    dT0.copy_(hT0, non_blocking=True); dcolp.copy_(hcolp, non_blocking=True)
    h.solve_dev(dT0, dcolp, d["glob"], HostOut(hout), na, nr, st, stream=s.cuda_stream)
    r = (na.cpu(), nr.cpu(), st.cpu())
    torch.cuda.synchronize()
    return r
ref = None
for name, fn in (("staged (library)", staged), ("zero copy", zero_copy), ("staged (library)", staged), ("zero copy", zero_copy)):
    fn(); fn()
    hout.zero_()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): fn()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
    chk = float(hout.double().sum())
    if ref is None: ref = chk
    print(f"{name}: {dt*1e3:.2f} ms per call, {1.08e9*B/65536/dt/1e9:.1f} GB/s of trajectories, checksum {chk:.6e} (same {chk == ref})", flush=True)
