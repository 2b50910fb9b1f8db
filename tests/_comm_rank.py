"""One rank of the C-ABI multi-GPU test (tests/test_gpu_comm.py): ctypes only -- no torch, no torch.distributed.
usage: python tests/_comm_rank.py RANK WORLD DEVICE IDFILE OUTFILE B"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nfc_b200
from oracle import nde_oracle as O

rank, world, device, idfile, outfile, B = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], sys.argv[5], int(sys.argv[6])
syn = nfc_b200.synthetic
d = syn.make_columns("ocean", B=B, seed=3)
cw, layers = O.chain_spec("dense-default")
theta = O.glorot_theta(cw, layers, 0, 0.3)
sv = syn.saveat(17)
Ttrue = np.repeat(d["T0"][:, None, :], 17, axis=1) * np.float32(0.97)
h = nfc_b200.Handle(cw, layers, 1, sv, device=device)
h.set_weights(theta)
if rank == 0:
    uid = nfc_b200.comm_unique_id()
    with open(idfile + ".tmp", "wb") as f:
        f.write(uid)
    os.replace(idfile + ".tmp", idfile)          # the host ships the id: here through the file system
else:
    t0 = time.time()
    while not os.path.exists(idfile):
        if time.time() - t0 > 60:
            raise SystemExit("rank 0 never wrote the rendezvous id")
        time.sleep(0.05)
    uid = open(idfile, "rb").read()
h.comm_init(uid, rank, world)
lo = rank * B // world + (0 if rank else 0)
hi = (rank + 1) * B // world
if len(sys.argv) > 7 and sys.argv[7] == "empty_last" and rank == world - 1:
    lo = hi = B                                   # a rank without columns still takes part in the collective
elif len(sys.argv) > 7 and sys.argv[7] == "empty_last":
    hi = B if rank == world - 2 else hi
r = h.loss_grad(d["T0"][lo:hi], d["colp"][lo:hi], d["glob"], Ttrue[lo:hi])
c = h.counters()
np.savez(outfile, loss=r["loss"], grad=r["grad"], status=r["status"], comm_world=c["comm_world"], allreduce_ms=c["last_allreduce_ms"])
h.comm_destroy()
