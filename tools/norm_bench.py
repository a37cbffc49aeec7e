"""Micro-benchmark of the HBM-bound helpers (GroupNorm, LayerNorm, concat/FreeU, upsample) on the step's shapes."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from instantrestore_b200 import _lib as L


def timeit(f, n=20):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def main():
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for (B, HW, C) in [(1, 262144, 128), (4, 262144, 128), (1, 65536, 256), (4, 65536, 256), (1, 16384, 512), (1, 4096, 512),
                       (4, 4096, 512), (1, 4096, 320), (4, 4096, 320), (1, 1024, 640), (1, 256, 1280), (1, 64, 1280), (8, 262144, 128)]:
        x = torch.randn(B * HW, C, device="cuda").half()
        g, b = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
        out = torch.empty_like(x)
        ws = torch.empty(L.load().ir_groupnorm_workspace_bytes(B, 32) // 4, dtype=torch.float32, device="cuda")
        us = timeit(lambda: L.groupnorm(x, g, b, batch=B, hw=HW, silu=True, out=out, workspace=ws))
        nbytes = 3 * x.numel() * 2
        print(f"groupnorm B={B} HW={HW:6d} C={C:4d}: {us:8.1f} us  {nbytes / us / 1e6:7.2f} TB/s (3 passes)", flush=True)
    for (R, C) in [(16384, 320), (4096, 640), (1024, 1280), (131072, 320)]:
        x = torch.randn(R, C, device="cuda").half()
        g, b = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
        out = torch.empty_like(x)
        us = timeit(lambda: L.layernorm(x, g, b, out=out))
        print(f"layernorm R={R:6d} C={C:4d}: {us:8.1f} us  {2 * x.numel() * 2 / us / 1e6:7.2f} TB/s (2 passes)", flush=True)
    a = torch.randn(1 << 28, device="cuda", dtype=torch.float16)
    bb = torch.empty_like(a)
    us = timeit(lambda: bb.copy_(a), 10)
    print(f"torch copy 512 MiB: {us:8.1f} us {2 * a.numel() * 2 / us / 1e6:7.2f} TB/s")


if __name__ == "__main__":
    main()
