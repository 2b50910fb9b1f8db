"""world_size-2 gloo test (CPU) of the N>1 host logic: column sharding + the single fused all-reduce of [grad; loss].
Each rank computes its shard with the CPU oracle; the combined result must equal the oracle on the whole batch."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import sys
    sys.path.insert(0, ROOT)
    import nfc_b200
    from oracle import c_oracle as CO
    from oracle import nde_oracle as O
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    syn = nfc_b200.synthetic
    B = 7                                                    # ragged: 4 + 3
    d = syn.make_columns("training9")
    cw, layers = O.chain_spec()
    theta = O.glorot_theta(cw, layers, 0, 1.0)
    sv = syn.saveat(9)
    T0, colp = d["T0"][:B], d["colp"][:B]
    Ttrue = np.repeat(T0[:, None, :], 9, axis=1) + 0.05
    lo, hi = nfc_b200.distributed.shard_range(B, rank, world)
    r = CO.loss_grad_continuous(T0[lo:hi], theta, cw, layers, colp[lo:hi], d["glob"], sv, Ttrue[lo:hi])
    nb = hi - lo
    # the oracle normalises by its own shard; the library normalises by the global count (n_total_columns)
    g_share = torch.from_numpy(r["grad"] * (nb / B))
    loss_sum = r["loss"] * nb * 9 * 32
    g, lsum = nfc_b200.distributed.fuse_and_allreduce(g_share, loss_sum)
    if rank == 0:
        full = CO.loss_grad_continuous(T0, theta, cw, layers, colp, d["glob"], sv, Ttrue)
        np.savez(os.path.join(out_dir, "res.npz"), g=g.numpy(), loss=lsum / (B * 9 * 32), g_full=full["grad"], loss_full=full["loss"],
                 shards=np.array([lo, hi]))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_partitions():
    import sys
    sys.path.insert(0, ROOT)
    import nfc_b200
    for B in (0, 1, 7, 8, 65536, 1048576 + 3):
        for world in (1, 2, 4, 8):
            rs = [nfc_b200.distributed.shard_range(B, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == B
            assert all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in rs]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_gradient_allreduce_matches_full_batch(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    res = np.load(tmp_path / "res.npz")
    assert abs(res["loss"] - res["loss_full"]) / res["loss_full"] < 1e-6
    assert np.linalg.norm(res["g"] - res["g_full"]) / np.linalg.norm(res["g_full"]) < 1e-5
