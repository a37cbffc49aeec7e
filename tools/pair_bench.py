"""A/B of the CTA-pair (tcgen05 cta_group::2) conv/GEMM kernel against the single-CTA kernels: equality of results
(same K order, so the outputs are expected to match bit for bit) and time per launch.
usage: python tools/pair_bench.py [check]"""
import math
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from instantrestore_b200 import _lib as L
from instantrestore_b200.weights import geglu_interleave_index

SHAPES = [  # (kind, batch, H, Cin, Cout); for lin/geglu H = tokens
    ("conv3", 1, 64, 512, 512), ("conv3", 8, 128, 512, 512), ("conv3", 8, 256, 256, 256), ("conv3", 2, 256, 256, 256),
    ("conv3", 8, 16, 1280, 1280), ("conv3", 32, 16, 1280, 1280), ("conv3", 8, 32, 640, 1280), ("conv3", 4, 128, 512, 512),
    ("conv3s2", 8, 256, 256, 256), ("conv3", 8, 512, 128, 128), ("conv3", 2, 512, 128, 128), ("conv3", 4, 256, 128, 128),
    ("conv3", 1, 128, 512, 512), ("conv3", 4, 64, 512, 512), ("conv3", 8, 64, 320, 640), ("conv3", 8, 256, 128, 256),
    ("conv3", 32, 64, 320, 320), ("conv3", 32, 32, 640, 640), ("conv3", 8, 64, 640, 320), ("conv3", 8, 32, 1280, 640), ("conv3", 4, 64, 320, 320),
    ("lin", 1, 131072, 320, 320), ("lin", 1, 32768, 640, 640), ("lin", 1, 131072, 1280, 320), ("lin", 1, 32768, 2560, 640),
    ("lin", 1, 8192, 1280, 1280), ("lin", 1, 4096, 512, 4096), ("lin", 1, 32768, 1280, 1280), ("lin", 1, 1000, 512, 512),
    ("geglu", 1, 131072, 320, 2560), ("geglu", 1, 32768, 640, 5120), ("geglu", 1, 8192, 1280, 10240),
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
    check_only = len(sys.argv) > 1 and sys.argv[1] == "check"
    g = torch.Generator(device="cuda").manual_seed(0)
    bad = 0
    for kind, B, H, Ci, Co in SHAPES:
        if kind.startswith("conv3"):
            st = 2 if kind.endswith("s2") else 1
            Ho = H // st
            a = torch.randn(B * H * H, Ci, device="cuda", generator=g).half()
            w = (torch.randn(Co, 9 * Ci, device="cuda", generator=g) / math.sqrt(9 * Ci)).half()
            bias = torch.randn(Co, device="cuda", generator=g)
            res = torch.randn(B * Ho * Ho, Co, device="cuda", generator=g).half()
            f = lambda cp: L.conv_gemm(a, w, batch=B, h_in=H, w_in=H, c_in=Ci, ksize=3, stride=st, bias=bias, residual=res, cta_pair=cp, split_k=(0 if cp == 0 else 1))
            flops = 2.0 * B * Ho * Ho * 9 * Ci * Co
        else:
            M = H
            a = torch.randn(M, Ci, device="cuda", generator=g).half()
            w = (torch.randn(Co, Ci, device="cuda", generator=g) / math.sqrt(Ci)).half()
            bias = torch.randn(Co, device="cuda", generator=g)
            if kind == "geglu":
                idx = geglu_interleave_index(Co).cuda()
                w, bias = w[idx].contiguous(), bias[idx].contiguous()
                f = lambda cp: L.conv_gemm(a, w, batch=1, h_in=1, w_in=M, c_in=Ci, bias=bias, act=L.IR_ACT_GEGLU, cta_pair=cp)
            else:
                res = torch.randn(M, Co, device="cuda", generator=g).half()
                f = lambda cp: L.conv_gemm(a, w, batch=1, h_in=1, w_in=M, c_in=Ci, bias=bias, residual=res, cta_pair=cp, split_k=(0 if cp == 0 else 1))
            flops = 2.0 * M * Ci * Co
        o1, o2 = f(1), f(2)
        o0 = f(0)
        torch.cuda.synchronize()
        diff = (o1.float() - o2.float()).abs().max().item()
        ref = o1.float().abs().max().item()
        ok = diff <= 2e-3 * max(ref, 1.0)
        bad += 0 if ok else 1
        line = f"{kind:7s} B={B:2d} H/M={H:7d} {Ci:5d}->{Co:5d}: max|pair - single| {diff:.3e} (max|out| {ref:.2f}) {'OK' if ok else 'MISMATCH'}"
        if not check_only:
            t1, t2, t0 = timeit(lambda: f(1)), timeit(lambda: f(2)), timeit(lambda: f(0))
            line += f" | single {t1:8.1f} us {flops / t1 / 1e6:7.1f} TF/s | pair {t2:8.1f} us {flops / t2 / 1e6:7.1f} TF/s | auto {flops / t0 / 1e6:7.1f}"
        print(line, flush=True)
    print("pair_bench:", "ALL OK" if bad == 0 else f"{bad} MISMATCHES")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
