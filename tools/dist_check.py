"""torchrun multi-rank check on real GPUs: sharded loss+gradient with ONE NCCL all-reduce equals the single-GPU full batch."""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nfc_b200

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
syn = nfc_b200.synthetic
B = 4099
d = syn.make_columns("ocean", B=B, seed=7)
NN = nfc_b200.named_chain("dense-default", seed=0)
theta, _ = nfc_b200.destructure(NN)
sv = syn.saveat(33)
Ttrue = np.repeat(d["T0"][:, None, :], 33, axis=1) * 0.98
h = nfc_b200.Handle(NN.conv_width, NN.dense_spec, 1, sv, device=local)
h.set_weights(theta)
lo, hi = nfc_b200.distributed.shard_range(B, rank, world)
T0 = torch.from_numpy(d["T0"][lo:hi]).to(dev); colp = torch.from_numpy(d["colp"][lo:hi]).to(dev); Tt = torch.from_numpy(Ttrue[lo:hi]).to(dev)
s = torch.cuda.Stream(device=dev); torch.cuda.set_stream(s)
loss, g, st = nfc_b200.distributed.loss_grad_sharded(h, T0, colp, d["glob"], Tt, n_total_columns=B, stream=s.cuda_stream)
torch.cuda.synchronize()
if rank == 0:
    full = h.loss_grad(d["T0"], d["colp"], d["glob"], Ttrue)
    rel = np.linalg.norm(g.cpu().numpy() - full["grad"]) / np.linalg.norm(full["grad"])
    print(f"world {world}: sharded loss {loss:.9e} full {full['loss']:.9e} grad rel diff {rel:.3e} status_bad {int((st != 0).sum())}", flush=True)
    assert abs(loss - full["loss"]) / full["loss"] < 1e-6 and rel < 2e-4
    print("DIST CHECK OK", flush=True)
dist.barrier()
dist.destroy_process_group()
