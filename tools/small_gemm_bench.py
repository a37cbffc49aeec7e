"""Device time per launch of the single-identity GEMM / conv shapes (the worst rows of bench.py's roofline), measured
like bench.py does: 8 back-to-back launches inside a CUDA graph, best of 5 replays. A-B a previous build with IR_LIB_PATH."""
import math
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from instantrestore_b200 import _lib as L

SHAPES = [  # (kind, M or (B, H), K / Cin, N / Cout, residual)
    ("lin", 4096, 640, 640, True), ("lin", 1024, 1280, 1280, True), ("lin", 256, 1280, 1280, True), ("lin", 1024, 640, 640, True),
    ("lin", 4096, 320, 320, True), ("lin", 16384, 320, 320, True), ("lin", 4096, 640, 1920, False), ("lin", 256, 1280, 3840, False),
    ("lin", 4096, 512, 512, True), ("lin", 4096, 512, 4096, False), ("lin", 4096, 4096, 512, False), ("lin", 4096, 2560, 640, True),
    ("lin", 1024, 5120, 1280, True), ("lin", 4096, 1920, 640, True), ("conv", (1, 64), 640, 640, True), ("conv", (1, 64), 960, 320, True),
    ("conv", (1, 16), 1280, 1280, True), ("conv", (1, 64), 512, 512, True), ("conv", (1, 32), 1280, 1280, True), ("conv", (1, 8), 1280, 1280, False),
]


def main():
    g = torch.Generator(device="cuda").manual_seed(0)
    kw = {}
    for a_ in sys.argv[1:]:             # e.g. split_k=1 tile_n=64
        k_, v_ = a_.split("=")
        kw[k_] = int(v_)
    for kind, m, k, n, res in SHAPES:
        if kind == "lin":
            a = torch.randn(m, k, device="cuda", generator=g).half()
            w = (torch.randn(n, k, device="cuda", generator=g) / math.sqrt(k)).half()
            r = torch.randn(m, n, device="cuda", generator=g).half() if res else None
            bias = torch.randn(n, device="cuda", generator=g)
            out = torch.empty(m, n, device="cuda", dtype=torch.float16)
            f = lambda: L.conv_gemm(a, w, batch=1, h_in=1, w_in=m, c_in=k, bias=bias, residual=r, out=out, **kw)
            flops, tag = 2.0 * m * k * n, f"lin  m{m}_k{k}_n{n}"
        else:
            b, h = m
            a = torch.randn(b * h * h, k, device="cuda", generator=g).half()
            w = (torch.randn(n, 9 * k, device="cuda", generator=g) / math.sqrt(9 * k)).half()
            r = torch.randn(b * h * h, n, device="cuda", generator=g).half() if res else None
            bias = torch.randn(n, device="cuda", generator=g)
            out = torch.empty(b * h * h, n, device="cuda", dtype=torch.float16)
            f = lambda: L.conv_gemm(a, w, batch=b, h_in=h, w_in=h, c_in=k, ksize=3, bias=bias, residual=r, out=out, **kw)
            flops, tag = 2.0 * b * h * h * 9 * k * n, f"conv m{b * h * h}_k{9 * k}_n{n}"
        try:
            f()
        except RuntimeError as e:
            print(f"{tag:28s} res={int(res)}  n/a ({str(e)[:60]})", flush=True)
            continue
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            for _ in range(8):
                f()
        gr.replay()
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            gr.replay()
            e1.record()
            e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        us = best * 1e3 / 8
        print(f"{tag:28s} res={int(res)}  {us:7.2f} us  {flops / us / 1e6:7.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    main()
