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


def broadcast_meta(sds, src: int = 0):
    """Key / shape / dtype of checkpoint state dicts from `src` (the only thing the other ranks need BEFORE the prepared
    weights arrive): returns, on every rank, a list of {key: (shape, dtype)} dicts."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [{k: (tuple(v.shape), v.dtype) for k, v in sd.items()} for sd in sds]
    box = [None]
    if dist.get_rank() == src:
        box = [[{k: (tuple(v.shape), str(v.dtype).replace("torch.", "")) for k, v in sd.items()} for sd in sds]]
    dist.broadcast_object_list(box, src=src)
    return [{k: (sh, getattr(torch, dt) if isinstance(dt, str) else dt) for k, (sh, dt) in m.items()} for m in box[0]]


def placeholder_state_dict(meta) -> Dict[str, torch.Tensor]:
    """Zero tensors with the checkpoint's keys / shapes / dtypes (calloc'ed pages: nothing is touched until the engine's
    load-time folding reads them). A rank that is about to RECEIVE prepared weights builds its engine from this."""
    return {k: torch.zeros(sh, dtype=dt) for k, (sh, dt) in meta.items()}


def engine_tensors(obj, device=None):
    """Every tensor an engine object tree holds, as (path, tensor) in a deterministic depth-first order (attribute /
    key insertion order): the packed fp16 weights, fp32 biases and norm parameters, folded constants. Objects of this
    package, lists, tuples and dicts are walked; anything else (streams, generators, CUDA graphs) is skipped."""
    out, seen = [], set()

    def walk(x, path):
        if isinstance(x, torch.Tensor):
            if id(x) not in seen and (device is None or x.device == torch.device(device)):
                seen.add(id(x))
                out.append((path, x))
            return
        if id(x) in seen:
            return
        if isinstance(x, dict):
            seen.add(id(x))
            for k, v in x.items():
                walk(v, f"{path}[{k!r}]")
        elif isinstance(x, (list, tuple)):
            seen.add(id(x))
            for i, v in enumerate(x):
                walk(v, f"{path}[{i}]")
        elif hasattr(x, "__dict__") and type(x).__module__.split(".")[0] in ("instantrestore_b200", "__main__", "tests", "test_dist_gloo"):
            seen.add(id(x))
            for k, v in vars(x).items():
                if k.startswith("_graphs") or k in ("_gen", "_side", "debug", "captured", "attention_probs", "reference_mass", "skip_acts"):
                    continue
                walk(v, f"{path}.{k}")

    walk(obj, type(obj).__name__)
    return out


def broadcast_engine(engine, src: int = 0, device=None) -> dict:
    """Replicates the PREPARED weights of `engine` (LoRA merged, folded, packed fp16 — what the kernels read) from rank
    `src` to the same engine structure on every other rank: the tensors are flattened into one arena per dtype on the
    device, ONE device-to-device broadcast per dtype (NCCL over NVLink; gloo in the CPU tests), and copied back into
    place on the receivers. This is the start-up broadcast of SURVEY.md 8e: ~3.9 GB for two UNets + two VAEs, tens of
    milliseconds, instead of shipping the raw fp32 checkpoint through host memory and re-preparing it on every rank.
    Returns {'tensors': n, 'bytes': total}."""
    items = engine_tensors(engine, device)
    total = sum(t.numel() * t.element_size() for _, t in items)
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return {"tensors": len(items), "bytes": total}
    sig = [(p, tuple(t.shape), str(t.dtype)) for p, t in items]
    box = [sig if dist.get_rank() == src else None]
    dist.broadcast_object_list(box, src=src)
    if box[0] != sig:
        diff = next((a, b) for a, b in zip(box[0] + [None] * len(sig), sig + [None] * len(box[0])) if a != b)
        raise RuntimeError(f"broadcast_engine: rank {dist.get_rank()} holds a different engine structure than rank {src}: {diff}")
    by_dtype: Dict[torch.dtype, list] = {}
    for _, t in items:
        by_dtype.setdefault(t.dtype, []).append(t)
    for dt, ts in by_dtype.items():
        dev = ts[0].device
        n = sum(t.numel() for t in ts)
        if dist.get_rank() == src:
            arena = torch.cat([t.reshape(-1) for t in ts])
        else:
            arena = torch.empty(n, dtype=dt, device=dev)
        dist.broadcast(arena, src=src)
        if dist.get_rank() != src:
            off = 0
            for t in ts:
                m = t.numel()
                t.copy_(arena[off:off + m].view(t.shape))
                off += m
        del arena
    return {"tensors": len(items), "bytes": total}


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
