"""Micro-benchmark of ir_conv_gemm on the step's dominant shapes: persistent kernel vs one-tile-per-CTA kernel."""
import math
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from instantrestore_b200 import _lib as L
from instantrestore_b200.weights import geglu_interleave_index

SHAPES = [  # (kind, batch, H, Cin, Cout)  kind: conv3 | lin | geglu ; for lin/geglu H = tokens
    ("conv3", 8, 512, 128, 128), ("conv3", 2, 512, 128, 128), ("conv3", 8, 256, 256, 256), ("conv3", 8, 128, 512, 512),
    ("conv3", 8, 64, 320, 320), ("conv3", 8, 32, 640, 640), ("conv3", 8, 16, 1280, 1280), ("conv3", 1, 64, 320, 320),
    ("geglu", 1, 131072, 320, 2560), ("geglu", 1, 32768, 640, 5120), ("geglu", 1, 8192, 1280, 10240), ("geglu", 1, 16384, 320, 2560),
    ("lin", 1, 131072, 320, 320), ("lin", 1, 32768, 640, 640), ("lin", 1, 8192, 1280, 1280), ("lin", 1, 131072, 1280, 320),
    ("lin", 1, 4096, 512, 4096), ("lin", 1, 16384, 320, 320),
]


def timeit(f, n=10):
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
    g = torch.Generator(device="cuda").manual_seed(0)
    shapes = SHAPES
    if len(sys.argv) >= 6:      # one shape from the command line: kind B H Cin Cout
        shapes = [(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]))]
    for kind, B, H, Ci, Co in shapes:
        if kind == "conv3":
            a = torch.randn(B * H * H, Ci, device="cuda", generator=g).half()
            w = (torch.randn(Co, 9 * Ci, device="cuda", generator=g) / math.sqrt(9 * Ci)).half()
            bias = torch.randn(Co, device="cuda", generator=g)
            res = torch.randn(B * H * H, Co, device="cuda", generator=g).half()
            out = torch.empty(B * H * H, Co, device="cuda", dtype=torch.float16)
            f = lambda np_: L.conv_gemm(a, w, batch=B, h_in=H, w_in=H, c_in=Ci, ksize=3, bias=bias, residual=res, out=out, no_persistent=np_)
            flops = 2.0 * B * H * H * 9 * Ci * Co
        else:
            M = H
            a = torch.randn(M, Ci, device="cuda", generator=g).half()
            w = (torch.randn(Co, Ci, device="cuda", generator=g) / math.sqrt(Ci)).half()
            bias = torch.randn(Co, device="cuda", generator=g)
            if kind == "geglu":
                idx = geglu_interleave_index(Co).cuda()
                w, bias = w[idx].contiguous(), bias[idx].contiguous()
                out = torch.empty(M, Co // 2, device="cuda", dtype=torch.float16)
                f = lambda np_: L.conv_gemm(a, w, batch=1, h_in=1, w_in=M, c_in=Ci, bias=bias, act=L.IR_ACT_GEGLU, out=out, no_persistent=np_)
            else:
                res = torch.randn(M, Co, device="cuda", generator=g).half()
                out = torch.empty(M, Co, device="cuda", dtype=torch.float16)
                f = lambda np_: L.conv_gemm(a, w, batch=1, h_in=1, w_in=M, c_in=Ci, bias=bias, residual=res, out=out, no_persistent=np_)
            flops = 2.0 * M * Ci * Co
        t_old, t_new = timeit(lambda: f(True)), timeit(lambda: f(False))
        print(f"{kind:6s} B={B:2d} H/M={H:7d} {Ci:5d}->{Co:5d}: one-tile {t_old:8.1f} us {flops / t_old / 1e6:7.1f} TF/s | persistent {t_new:8.1f} us {flops / t_new / 1e6:7.1f} TF/s", flush=True)


if __name__ == "__main__":
    main()
