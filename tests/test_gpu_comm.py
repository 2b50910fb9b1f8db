"""Multi-GPU behind the C ABI (include/nfc.h: nfc_comm_unique_id / nfc_comm_init; SURVEY.md 8b "with G>1 GPUs performs the allreduce
internally", 8e): every rank calls nfc_loss_grad on its own columns and gets the loss and gradient of the GLOBAL batch
(src/training.jl:45-48,70 over all simulations).  The ranks are plain python processes that talk to libnfc_b200.so through ctypes and
exchange the NCCL rendezvous id through a file: no torch.distributed anywhere."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import nde_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _rel(a, b):
    return float(np.linalg.norm(np.asarray(a, np.float64) - np.asarray(b, np.float64)) / np.linalg.norm(np.asarray(b, np.float64)))


def _single(nfc, B):
    syn = nfc.synthetic
    d = syn.make_columns("ocean", B=B, seed=3)
    cw, layers = O.chain_spec("dense-default")
    theta = O.glorot_theta(cw, layers, 0, 0.3)
    sv = syn.saveat(17)
    Ttrue = np.repeat(d["T0"][:, None, :], 17, axis=1) * np.float32(0.97)
    h = nfc.Handle(cw, layers, 1, sv)
    h.set_weights(theta)
    return h, d, Ttrue, h.loss_grad(d["T0"], d["colp"], d["glob"], Ttrue)


def _ranks(tmp_path, world, devices, B, mode=""):
    idfile = str(tmp_path / "nccl_id")
    procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_comm_rank.py"), str(r), str(world), str(devices[r]), idfile,
                               str(tmp_path / f"rank{r}.npz"), str(B)] + ([mode] if mode else []), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(world)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
    return [dict(np.load(tmp_path / f"rank{r}.npz")) for r in range(world)]


def test_world_of_one_goes_through_the_collective(nfc, gpu_ok, tmp_path):
    h, d, Ttrue, ref = _single(nfc, 64)
    (r,) = _ranks(tmp_path, 1, [0], 64)
    assert int(r["comm_world"]) == 1
    assert abs(float(r["loss"]) - ref["loss"]) <= 1e-12 * ref["loss"]
    assert _rel(r["grad"], ref["grad"]) < 1e-6           # same kernels; the normalisation happens after the all-reduce instead of before


@pytest.mark.parametrize("mode", ["", "empty_last"])
def test_two_ranks_through_the_c_abi_only(nfc, gpu_ok, tmp_path, mode):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2); the one-rank communicator test covers the code path on one GPU")
    h, d, Ttrue, ref = _single(nfc, 96)
    r0, r1 = _ranks(tmp_path, 2, [0, 1], 96, mode)
    for r in (r0, r1):
        assert int(r["comm_world"]) == 2
        assert abs(float(r["loss"]) - ref["loss"]) <= 1e-9 * ref["loss"]
        assert _rel(r["grad"], ref["grad"]) < 1e-5       # sum over two ranks vs one: float32 summation order only
    assert np.array_equal(r0["grad"], r1["grad"])        # every rank holds the same all-reduced gradient
