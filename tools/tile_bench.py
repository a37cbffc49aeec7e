"""A/B of the GEMM tiling choices (N-tile width, persistent vs one-tile kernel, stacked M tiles) on 3x3 convolutions.
usage: python tools/tile_bench.py            (prints TFLOP/s per variant and shape)"""
import math
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from instantrestore_b200 import _lib as L  # noqa: E402
from tools.gemm_bench import timeit  # noqa: E402

SHAPES = [  # (batch, H, Cin, Cout)
    (8, 512, 128, 128), (1, 512, 128, 128), (8, 256, 256, 256), (1, 256, 256, 256), (8, 128, 512, 512), (1, 128, 512, 512),
    (8, 64, 512, 512), (8, 64, 320, 320), (8, 32, 640, 640), (8, 32, 1280, 640), (8, 16, 1280, 1280), (8, 16, 2560, 1280),
]


def main():
    g = torch.Generator(device="cuda").manual_seed(0)
    for B, H, Ci, Co in SHAPES:
        a = torch.randn(B * H * H, Ci, device="cuda", generator=g).half()
        w = (torch.randn(Co, 9 * Ci, device="cuda", generator=g) / math.sqrt(9 * Ci)).half()
        bias = torch.randn(Co, device="cuda", generator=g)
        out = torch.empty(B * H * H, Co, device="cuda", dtype=torch.float16)
        flops = 2.0 * B * H * H * 9 * Ci * Co
        res = []
        t = timeit(lambda: L.conv_gemm(a, w, batch=B, h_in=H, w_in=H, c_in=Ci, ksize=3, bias=bias, out=out))
        res.append(f"auto {flops / t / 1e6:6.0f}")
        t = timeit(lambda: L.conv_gemm(a, w, batch=B, h_in=H, w_in=H, c_in=Ci, ksize=3, bias=bias, out=out, m_sub=1))
        res.append(f"auto/no-stack {flops / t / 1e6:6.0f}")
        for tn in (128, 160, 256):
            if (tn == 256 and Co % 256) or (tn == 160 and Co % 160):
                continue
            for npers, tag in ((1, "1tile"), (2, "pers")):
                t = timeit(lambda: L.conv_gemm(a, w, batch=B, h_in=H, w_in=H, c_in=Ci, ksize=3, bias=bias, out=out, tile_n=tn,
                                               no_persistent=npers, split_k=1, m_sub=1))
                res.append(f"tn{tn}/{tag} {flops / t / 1e6:6.0f}")
        print(f"B={B} H={H:3d} {Ci:4d}->{Co:4d}: " + " | ".join(res), flush=True)


if __name__ == "__main__":
    main()
