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
