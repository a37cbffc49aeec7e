"""A/B of N-tile width and kernel variant for the long-K convolutions."""
import math, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from instantrestore_b200 import _lib as L
from tools.gemm_bench import timeit

SHAPES = [(8, 512, 128, 128), (1, 512, 128, 128), (4, 512, 128, 128), (8, 256, 128, 128), (1, 256, 128, 128), (8, 512, 64, 128)]
_OLD2 = [(8, 16, 1280, 1280), (8, 16, 2560, 1280), (8, 32, 1280, 1280), (8, 8, 1280, 1280), (4, 16, 1280, 1280), (8, 32, 640, 640), (8, 32, 1280, 640), (8, 32, 1920, 640), (8, 64, 640, 320), (8, 64, 960, 320), (1, 32, 640, 640), (4, 32, 640, 640)]
_OLD = [(8, 128, 512, 512), (1, 128, 512, 512), (8, 256, 256, 256), (1, 256, 256, 256), (8, 32, 640, 640), (8, 16, 1280, 1280),
          (8, 64, 320, 320), (8, 32, 1280, 640), (4, 64, 512, 512), (1, 64, 512, 512), (8, 64, 512, 512), (8, 16, 2560, 1280), (8, 32, 1920, 640),
          (8, 256, 128, 256), (8, 128, 256, 512), (8, 512, 128, 128)]
g = torch.Generator(device="cuda").manual_seed(0)
for B, H, Ci, Co in SHAPES:
    a = torch.randn(B * H * H, Ci, device="cuda", generator=g).half()
    w = (torch.randn(Co, 9 * Ci, device="cuda", generator=g) / math.sqrt(9 * Ci)).half()
    bias = torch.randn(Co, device="cuda", generator=g)
    out = torch.empty(B * H * H, Co, device="cuda", dtype=torch.float16)
    flops = 2.0 * B * H * H * 9 * Ci * Co
    res = []
    for msub in (1, 0):
        t = timeit(lambda: L.conv_gemm(a, w, batch=B, h_in=H, w_in=H, c_in=Ci, ksize=3, bias=bias, out=out, m_sub=msub))
        res.append(f"auto m_sub={msub} {flops / t / 1e6:6.0f}")
    for tn in ():
        for npers in (1, 2):
            if (tn == 256 and Co % 256) or (tn == 160 and Co % 160):
                continue
            try:
                f = lambda: L.conv_gemm(a, w, batch=B, h_in=H, w_in=H, c_in=Ci, ksize=3, bias=bias, out=out, tile_n=tn, no_persistent=npers, split_k=1)
                t = timeit(f)
                res.append(f"tn={tn:3d}{'/1tile' if npers == 1 else '/pers '} {flops / t / 1e6:6.0f}")
            except Exception as e:  # noqa
                res.append(f"tn={tn} err")
    print(f"B={B} H={H:3d} {Ci:4d}->{Co:4d}: " + " | ".join(res), flush=True)
