"""Process-group helpers: one process per GPU, NCCL on GPUs / gloo on CPU (tests).

The reference has no collective backend at all (its "distributed" layer is WAN libp2p, reference
infomesh/p2p/node.py:535-597); this module is the intra-node plumbing for the sharded index and the
tensor-parallel models.  Everything degenerates to no-ops at world size 1.
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import torch
import torch.distributed as dist


@dataclass
class DistContext:
    rank: int = 0
    world: int = 1
    local_rank: int = 0
    device: torch.device = torch.device("cpu")
    backend: str = "none"

    @property
    def is_dist(self) -> bool:
        return self.world > 1


_ctx: DistContext | None = None


def init(backend: str | None = None) -> DistContext:
    """Initialise from torchrun-style env (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*)."""
    global _ctx
    if _ctx is not None:
        return _ctx
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", str(rank)))
    cuda = torch.cuda.is_available()
    device = torch.device(f"cuda:{local_rank}") if cuda else torch.device("cpu")
    if cuda:
        torch.cuda.set_device(device)
    be = "none"
    if world > 1:
        be = backend or ("nccl" if cuda else "gloo")
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        if not dist.is_initialized():
            kwargs = {}
            if be == "nccl":
                kwargs["device_id"] = device
            dist.init_process_group(backend=be, rank=rank, world_size=world, **kwargs)
    _ctx = DistContext(rank, world, local_rank, device, be)
    return _ctx


def ctx() -> DistContext:
    return _ctx if _ctx is not None else init()


def shutdown(grace_s: float = 15.0) -> None:
    """Tear the process group down.  NCCL communicators that were captured into CUDA graphs can make
    ``destroy_process_group`` block forever; the teardown therefore runs under a watchdog and the process leaves
    with ``os._exit(0)`` (after flushing stdio) if it has not finished within ``grace_s``."""
    global _ctx
    if dist.is_initialized():
        import os
        import sys
        import threading

        if torch.cuda.is_available():
            torch.cuda.synchronize()
        try:
            dist.barrier()
        except Exception:  # noqa: BLE001
            pass
        done = threading.Event()

        def _destroy():
            try:
                dist.destroy_process_group()
            finally:
                done.set()

        threading.Thread(target=_destroy, daemon=True, name="pg-destroy").start()
        if not done.wait(grace_s):
            sys.stdout.flush()
            sys.stderr.flush()
            os._exit(0)
    _ctx = None


def barrier() -> None:
    if dist.is_initialized():
        dist.barrier()


def all_gather_cat(t: torch.Tensor) -> torch.Tensor:
    """All-gather equally-shaped tensors along a new leading dim -> [world, *t.shape]."""
    c = ctx()
    if not c.is_dist:
        return t.unsqueeze(0)
    out = torch.empty((c.world, *t.shape), device=t.device, dtype=t.dtype)
    dist.all_gather_into_tensor(out, t.contiguous())
    return out


def all_reduce_max(value: float) -> float:
    c = ctx()
    if not c.is_dist:
        return value
    t = torch.tensor([value], device=c.device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def all_reduce_sum_(t: torch.Tensor) -> torch.Tensor:
    if ctx().is_dist:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def broadcast_(t: torch.Tensor, src: int = 0) -> torch.Tensor:
    if ctx().is_dist:
        dist.broadcast(t, src=src)
    return t
