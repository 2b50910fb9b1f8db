"""Multi-GPU plumbing: columns are independent, so they shard over ranks with no data-path collective;
training needs ONE all-reduce per step of the small vector [dL/dtheta ; loss sum] (24 737 floats ~ 99 KB, latency bound
on NVLink 5 / NVSwitch).  One process per GPU, torch.distributed for the plumbing (backend nccl on GPUs, gloo in CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(B: int, rank: int, world: int):
    """Contiguous block of columns owned by `rank` (the first B % world ranks own one more)."""
    base, rem = divmod(B, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def fuse_and_allreduce(grad_local: torch.Tensor, loss_sum_local, group=None):
    """grad_local: float32 [n_params], this rank's share of d(loss)/dtheta (already divided by the GLOBAL residual count);
    loss_sum_local: this rank's sum of squared residuals (python float or 0-d float64 tensor).
    One all-reduce of a float32 buffer [n_params + 2]; the double loss sum travels as a (hi, lo) float pair.
    Returns (grad_global float32 [n_params], loss_sum_global float)."""
    n = grad_local.numel()
    buf = torch.empty(n + 2, dtype=torch.float32, device=grad_local.device)
    buf[:n] = grad_local
    ls = torch.as_tensor(loss_sum_local, dtype=torch.float64, device=grad_local.device).reshape(())
    hi = ls.to(torch.float32)
    buf[n] = hi
    buf[n + 1] = (ls - hi.to(torch.float64)).to(torch.float32)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return buf[:n], float(buf[n].double() + buf[n + 1].double())


CUDA_STREAM_LEGACY = 1          # cudaStreamLegacy: the NULL stream by its explicit handle (NULL itself means "the handle's own stream" in nfc.h)


def torch_stream_handle(dev):
    """The cudaStream_t of torch's current stream on `dev`, usable as the `stream` argument of the *_dev calls.  torch's default
    stream is the legacy NULL stream (handle 0): passing 0 would select the handle's own non-blocking stream, which is NOT ordered
    with work queued on torch's stream (zero fills of outputs, producers of the inputs, the previous step's update) -- so the legacy
    stream is passed by its explicit handle instead."""
    s = torch.cuda.current_stream(dev).cuda_stream
    return s if s else CUDA_STREAM_LEGACY


def attach_communicator(handle, group=None):
    """Gives `handle` its NCCL communicator (include/nfc.h: nfc_comm_init).  torch.distributed only ships the 128-byte rendezvous
    id from rank 0 to the others; the collective of a training step then runs inside nfc_loss_grad[_dev]."""
    from ._binding import comm_unique_id
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    ids = [comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0, group=group)
    handle.comm_init(ids[0], rank, world)
    handle.comm_world = world
    return handle


def loss_grad_sharded(handle, T0, colp, glob, Ttrue, n_total_columns: int, group=None, stream=None):
    """Loss and gradient of the GLOBAL batch; every rank passes its own shard as device tensors on handle.device.
    Mirrors nde_loss + Zygote.gradient (src/training.jl:45-48,70) over all ranks.  Returns (loss, grad, status) on every rank.
    A handle with a communicator (attach_communicator) all-reduces inside the C library; otherwise the all-reduce runs here on
    torch.distributed (the gloo CPU tests, or a host that wants its own process group)."""
    dev = T0.device
    n = handle.n_params
    grad = torch.empty(n, dtype=torch.float32, device=dev)
    ls = torch.zeros(2, dtype=torch.float64, device=dev)
    status = torch.zeros(max(T0.shape[0], 1), dtype=torch.int32, device=dev)
    if stream is None:
        stream = torch_stream_handle(dev)
    handle.loss_grad_dev(T0, colp, glob, Ttrue, ls, grad, status, n_total_columns=n_total_columns, stream=stream)
    if getattr(handle, "comm_world", 1) > 1:
        lsum = ls.cpu()                                  # synchronises the stream; ls = (global sum, global count)
        return float(lsum[0] / lsum[1]), grad, status
    g, lsum = fuse_and_allreduce(grad, ls[0], group)
    return lsum / (float(n_total_columns) * handle.n_save * handle.Nz), g, status
