"""Multi-GPU plumbing: one process per GPU, identities batch-sharded across ranks, weights replicated by ONE
broadcast at start-up (NCCL over NVLink/NVSwitch on the GPU box, gloo in CPU tests). There is no collective on the
hot path: every identity (degraded image + its references) is independent end to end."""
from __future__ import annotations

import os
from typing import Dict, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [lo, hi) slice of `n_items` identities owned by `rank`."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def init_from_env(backend: str | None = None) -> Tuple[int, int, int]:
    """(rank, world, local_rank) from torchrun's environment; initialises the default process group when world > 1."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def broadcast_state_dict(sd: Dict[str, torch.Tensor] | None, src: int = 0, device=None) -> Dict[str, torch.Tensor]:
    """Replicates a checkpoint from `src` to every rank with one metadata exchange and ONE flat-buffer broadcast per
    dtype (the 'trivial broadcast of shared UNet weights at startup'). Returns CPU tensors on every rank."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        assert sd is not None
        return sd
    rank = dist.get_rank()
    meta = [[(k, tuple(v.shape), str(v.dtype).replace("torch.", "")) for k, v in sd.items()]] if rank == src else [None]
    dist.broadcast_object_list(meta, src=src)
    meta = meta[0]
    dev = torch.device(device) if device is not None else (
        torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu"))
    out: Dict[str, torch.Tensor] = {}
    by_dtype: Dict[str, list] = {}
    for k, shape, dt in meta:
        by_dtype.setdefault(dt, []).append((k, shape))
    for dt, items in by_dtype.items():
        tdt = getattr(torch, dt)
        total = sum(int(torch.Size(s).numel()) for _, s in items)
        if rank == src:
            flat = torch.cat([sd[k].reshape(-1) for k, _ in items]).to(dev)
        else:
            flat = torch.empty(total, dtype=tdt, device=dev)
        dist.broadcast(flat, src=src)
        flat = flat.cpu()
        off = 0
        for k, shape in items:
            n = int(torch.Size(shape).numel())
            out[k] = flat[off:off + n].view(shape)
            off += n
    return {k: out[k] for k, _, _ in meta}


def max_over_ranks(value: float, device=None) -> float:
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    t = torch.tensor([value], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier() -> None:
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
