"""Micro-benchmark of ir_shared_attn_fwd on the layer shapes of the step (graph-replayed, CUDA events)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from instantrestore_b200 import _lib as L

SHAPES = [  # (B, H, S, own, n_ref, adain)
    (32, 5, 4096, True, 0, False), (8, 5, 4096, False, 4, True), (8, 5, 4096, True, 4, True), (32, 10, 1024, True, 0, False),
    (8, 10, 1024, False, 4, True), (4, 5, 4096, True, 0, False), (1, 5, 4096, False, 4, True), (1, 10, 1024, False, 4, True),
    (1, 20, 256, False, 4, True), (64, 5, 4096, False, 1, True), (4, 5, 4096, False, 8, True),
    (8, 5, 4096, False, 4, False), (64, 5, 4096, False, 1, False),   # the reference chunks through the plain variant
]


def main():
    g = torch.Generator(device="cuda").manual_seed(0)
    for (B, H, S, own, n_ref, adain) in SHAPES:
        C = H * 64
        q = torch.randn(B * S, 3 * C, device="cuda", generator=g).half()
        kw = {}
        s_kv = 0
        if own:
            kw.update(k_own=q[:, C:], v_own=q[:, 2 * C:], s_own=S); s_kv += S
        if n_ref:
            kv = torch.randn(B * n_ref * S, 3 * C, device="cuda", generator=g).half()
            kw.update(k_ref=kv[:, C:], v_ref=kv[:, 2 * C:], n_ref=n_ref, s_ref=S); s_kv += n_ref * S
            if adain:
                kw.update(adain_scale=torch.rand(B, n_ref, C, device="cuda") + 0.5, adain_shift=torch.randn(B, n_ref, C, device="cuda"))
        out = torch.empty(B * S, C, device="cuda", dtype=torch.float16)
        f = lambda: L.shared_attn(q, heads=H, scale=0.125, batch=B, s_q=S, out=out, **kw)
        for _ in range(3):
            f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 10
        e0.record()
        for _ in range(n):
            f()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        fl = 4.0 * B * H * S * s_kv * 64
        print(f"B={B:3d} H={H:2d} S={S:5d} own={int(own)} n_ref={n_ref} adain={int(adain)}  {ms * 1e3:9.1f} us  {fl / ms / 1e9:7.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    main()
