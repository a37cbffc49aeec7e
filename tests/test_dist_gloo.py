"""CPU, world_size 2 over gloo: the multi-GPU host logic (rank-0 weight broadcast, identity sharding,
max-over-ranks timing reduction)."""
import os
import socket
import sys
from pathlib import Path

import torch
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch.distributed as dist
    from instantrestore_b200 import dist as D
    from instantrestore_b200.synthetic import synthetic_unet_state_dict
    from instantrestore_b200.unet_engine import UNetSpec
    r, w, _ = D.init_from_env(backend="gloo")
    spec = UNetSpec(block_out_channels=(64, 128, 256, 256), attention_head_dim=(1, 2, 4, 4), cross_attention_dim=128)
    sd = synthetic_unet_state_dict(spec, seed=5, lora_rank=4) if r == 0 else None
    got = D.broadcast_state_dict(sd, src=0)
    want = synthetic_unet_state_dict(spec, seed=5, lora_rank=4)
    same = list(got.keys()) == list(want.keys()) and all(torch.equal(got[k], want[k]) for k in want)
    lo, hi = D.shard_range(7, r, w)
    t = D.max_over_ranks(1.0 + r)
    D.barrier()
    q.put((r, same, (lo, hi), t))
    dist.destroy_process_group()


def test_broadcast_and_sharding_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[1] for r in res] == [True, True]
    assert res[0][2] == (0, 4) and res[1][2] == (4, 7)
    assert res[0][3] == 2.0 and res[1][3] == 2.0


class _Leaf:          # stand-ins for the engine's _Lin / _Conv / _Norm holders (module name 'test_dist_gloo' is walked)
    def __init__(self, w, b):
        self.w, self.b = w, b


class _Tree:
    def __init__(self, seed):
        g = torch.Generator().manual_seed(seed)
        mk = lambda *s, dt=torch.float16: torch.randn(*s, generator=g).to(dt) if seed else torch.zeros(*s, dtype=dt)
        self.conv_in = _Leaf(mk(8, 36), mk(8, dt=torch.float32))
        self.blocks = [dict(norm=_Leaf(mk(8, dt=torch.float32), mk(8, dt=torch.float32)), lin=_Leaf(mk(16, 8), None)) for _ in range(3)]
        self.shared = self.blocks[0]["lin"].w            # an alias must be visited once
        self.scale = 0.5                                  # non-tensor attributes are skipped
        self._graphs = {"x": mk(4)}                       # run-time state is not part of the weights


def _engine_worker(rank, world, port, q):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch.distributed as dist
    from instantrestore_b200 import dist as D
    r, w, _ = D.init_from_env(backend="gloo")
    tree = _Tree(seed=11 if r == 0 else 0)               # rank 1 builds the same structure from placeholders
    info = D.broadcast_engine(tree, src=0)
    want = _Tree(seed=11)
    got, ref = D.engine_tensors(tree), D.engine_tensors(want)
    same = [p for p, _ in got] == [p for p, _ in ref] and all(torch.equal(a, b) for (_, a), (_, b) in zip(got, ref))
    meta = D.broadcast_meta([{"a.weight": torch.zeros(3, 4, dtype=torch.bfloat16), "b": torch.zeros(2)}] if r == 0 else None)
    ph = D.placeholder_state_dict(meta[0])
    ok_meta = ph["a.weight"].shape == (3, 4) and ph["a.weight"].dtype == torch.bfloat16 and ph["b"].dtype == torch.float32
    q.put((r, same, info["tensors"], ok_meta))
    dist.destroy_process_group()


def test_prepared_weight_broadcast_world2():
    """broadcast_engine: the prepared tensors of an engine object tree travel as one arena per dtype; receivers that built
    the same structure from placeholder weights end up bit-identical to rank 0."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_engine_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[1] for r in res] == [True, True]
    assert res[0][2] == res[1][2] == 2 + 3 * 3           # conv_in (w, b) + 3 x (norm w, norm b, lin w); alias and _graphs skipped
    assert [r[3] for r in res] == [True, True]
