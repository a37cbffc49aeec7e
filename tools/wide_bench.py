"""A/B of 256-bit epilogue loads / stores (ir_conv_gemm_params.wide_io) on the short-K layers whose epilogue keeps the
L1TEX LSU pipe busiest. Outputs must be bit-identical. usage: python tools/wide_bench.py"""
import math
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from instantrestore_b200 import _lib as L
from tools.halo_bench import timeit

SHAPES = [(2, 512, 128, 128), (8, 512, 128, 128), (4, 256, 128, 128), (4, 128, 512, 512), (32, 64, 320, 320), (8, 256, 256, 256)]


def main():
    g = torch.Generator(device="cuda").manual_seed(0)
    bad = 0
    for B, H, Ci, Co in SHAPES:
        a = torch.randn(B * H * H, Ci, device="cuda", generator=g).half()
        w = (torch.randn(Co, 9 * Ci, device="cuda", generator=g) / math.sqrt(9 * Ci)).half()
        bias = torch.randn(Co, device="cuda", generator=g)
        res = torch.randn(B * H * H, Co, device="cuda", generator=g).half()
        f = lambda wide: L.conv_gemm(a, w, batch=B, h_in=H, w_in=H, c_in=Ci, ksize=3, bias=bias, residual=res, wide_io=wide)
        same = torch.equal(f(1), f(0))
        bad += 0 if same else 1
        flops = 2.0 * B * H * H * 9 * Ci * Co
        t0, t1 = timeit(lambda: f(1)), timeit(lambda: f(0))        # wide_io: 1 = 128-bit only, 0 = auto (256-bit)
        print(f"conv3 B={B:2d} H={H:4d} {Ci:4d}->{Co:4d}: {'identical' if same else 'MISMATCH'} | 128-bit {t0:8.1f} us {flops / t0 / 1e6:7.1f} TF/s | "
              f"256-bit {t1:8.1f} us {flops / t1 / 1e6:7.1f} TF/s", flush=True)
    print("wide_bench:", "ALL OK" if bad == 0 else f"{bad} MISMATCHES")


if __name__ == "__main__":
    main()
