"""Bin-row sharding of a template stack across GPUs, one process per GPU.

GPU g holds rows [b_g, b_{g+1}) of every template and of the data; m_i needs only row i, so the composite,
the Poisson term and the residual are local; logL = sum_g logL_g and G = sum_g M_g' r_g.  The only exchange is ONE
all-reduce (sum, FP64) of [logL, G_1..G_T] per evaluation.  The `logL == 0 -> -Inf` guard (fitting_base.jl:95) and
fg!'s sign flip (solvers.jl:30-31) are applied AFTER the reduction.

Two equivalent ways to do the exchange:
  * in the library, on the kernel's stream:  init_library_comm(ctx, group)  ->  sfh_comm_init (NCCL)
  * in the host runtime:                     allreduce_fg(out, group)       ->  torch.distributed.all_reduce
The second also runs on the `gloo` backend and is what the CPU tests cover (world_size 2).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib as L


def shard_rows(nbins: int, world: int, rank: int, align: int = 32):
    """Contiguous balanced row ranges; interior boundaries aligned to `align` bins (tile-friendly).
    The union over ranks is exactly [0, nbins) for any nbins/world (shards may be empty)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    per = -(-nbins // world)
    per = -(-per // align) * align
    b = min(nbins, rank * per)
    e = min(nbins, (rank + 1) * per)
    return b, e


def guard_neg_logl(logl_raw: float) -> float:
    """fitting_base.jl:95 followed by solvers.jl:31: logL == 0 -> -typemax; return -logL."""
    return -logl_raw if logl_raw != 0.0 else float("inf")


def allreduce_fg(out, group=None):
    """out = [logL_raw, G...] of this rank's shard (torch tensor on any device, or numpy array).
    Sums it over the process group in place and returns (-logL guarded, G)."""
    import torch
    import torch.distributed as dist
    t = out if isinstance(out, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(out))
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    h = t.detach().cpu().numpy()
    return guard_neg_logl(float(h[0])), h[1:].copy()


def broadcast_bytes(payload: bytes | None, nbytes: int, src: int = 0, group=None) -> bytes:
    """Broadcast a fixed-size byte string (the 128-byte NCCL unique id) from `src` to every rank."""
    import torch
    import torch.distributed as dist
    buf = torch.zeros(nbytes, dtype=torch.uint8)
    if dist.get_rank(group) == src:
        buf = torch.frombuffer(bytearray(payload), dtype=torch.uint8).clone()
    dev = None
    if dist.get_backend(group) == "nccl":
        dev = torch.device("cuda", torch.cuda.current_device())
        buf = buf.to(dev)
    dist.broadcast(buf, src=src, group=group)
    return bytes(buf.cpu().numpy().tobytes())


def init_library_comm(ctx, group=None, p2p=True):
    """Give `ctx` (a fitting._Ctx) an NCCL communicator spanning the torch process group, so that every
    evaluation on it returns the all-reduced full-stack answer."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    idbuf = (C.c_char * 128)()
    if rank == 0:
        L.check(L.lib.sfh_comm_unique_id(idbuf))
    uid = broadcast_bytes(bytes(idbuf) if rank == 0 else None, 128, 0, group)
    L.check(L.lib.sfh_comm_init(ctx.handle, world, rank, uid))
    if p2p and world > 1 and os.environ.get("SFH_NO_P2P") != "1":
        init_p2p(ctx, group)
    return ctx


def all_ranks_agree(ok: bool, group=None) -> bool:
    """MIN over the group of a per-rank success flag: a collective every rank must call."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([1 if ok else 0], dtype=torch.int32)
    if dist.get_backend(group) == "nccl":
        t = t.to(torch.device("cuda", torch.cuda.current_device()))
    dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    return bool(int(t.cpu().item()) == 1)


def init_p2p(ctx, group=None, open_handles=None):
    """Exchange the CUDA-IPC handles of the per-rank inboxes and switch the fused path to the one-shot NVLink
    all-reduce inside the finalize kernel (include/sfhcuda.h: sfh_comm_p2p_*).  The switch is all-or-nothing: a rank
    whose peers' inboxes cannot be opened (no peer access in this topology, IPC disabled in its container) must not
    leave the others spinning on epoch flags it will never write.  So every rank tries, the group takes the MIN of the
    success flags, and unless every rank succeeded the ranks that did switch it off again (sfh_comm_p2p_enable(ctx, 0))
    and the NCCL all-reduce stays in place everywhere (returns False).  No evaluation runs in between.
    `open_handles(rank, blob) -> None | raises` replaces the library calls in the CPU (gloo) tests."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    h = (C.c_char * 64)()
    ok = True
    try:
        if open_handles is None:
            L.check(L.lib.sfh_comm_p2p_handle(ctx.handle, world, h))
    except (L.SFHError, ValueError):
        ok = False
    mine = torch.frombuffer(bytearray(bytes(h)), dtype=torch.uint8).clone()
    if dist.get_backend(group) == "nccl":
        mine = mine.to(torch.device("cuda", torch.cuda.current_device()))
    allh = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allh, mine, group=group)
    blob = b"".join(bytes(t.cpu().numpy().tobytes()) for t in allh)
    if ok:
        try:
            if open_handles is None:
                L.check(L.lib.sfh_comm_p2p_init(ctx.handle, world, rank, blob))
            else:
                open_handles(rank, blob)
        except (L.SFHError, ValueError):
            ok = False
    if all_ranks_agree(ok, group):
        return True
    if ok and open_handles is None:
        L.check(L.lib.sfh_comm_p2p_enable(ctx.handle, 0))
    return False
