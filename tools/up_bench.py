"""A/B of Upsample2D (nearest 2x + 3x3 conv): the materialised path (upsample kernel, then the 3x3 conv on the large
tensor) against ir_conv_gemm(upsample2x=1) = four 2x2 sub-pixel convolutions on the low-resolution input.
usage: python tools/up_bench.py [check]"""
import math
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from instantrestore_b200 import _lib as L
from instantrestore_b200.weights import upsample_conv_weight

SHAPES = [  # (batch, H low-res, C, want GroupNorm pass A)  — the step's upsamplers at B = 1 / 4 references / B = 8
    (1, 8, 1280, False), (4, 8, 1280, False), (8, 8, 1280, False), (32, 8, 1280, False),
    (1, 16, 1280, False), (4, 16, 1280, False), (8, 16, 1280, False), (32, 16, 1280, False),
    (1, 32, 640, False), (4, 32, 640, False), (8, 32, 640, False), (32, 32, 640, False),
    (1, 64, 512, True), (8, 64, 512, True), (1, 128, 512, True), (8, 128, 512, True), (1, 256, 256, True), (8, 256, 256, True),
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
    for B, H, Cc, stats in SHAPES:
        a = torch.randn(B * H * H, Cc, device="cuda", generator=g).half()
        w4 = (torch.randn(Cc, Cc, 3, 3, device="cuda", generator=g) / math.sqrt(9 * Cc)).half()
        wk = w4.permute(0, 2, 3, 1).contiguous().reshape(Cc, 9 * Cc)
        w_up = upsample_conv_weight(w4)
        bias = torch.randn(Cc, device="cuda", generator=g)
        part = torch.empty(L.gn_partial_numel(B, 4 * H * H), device="cuda") if stats else None
        big = lambda: L.conv_gemm(L.upsample_nearest2x(a, batch=B, h=H, w=H), wk, batch=B, h_in=2 * H, w_in=2 * H, c_in=Cc, ksize=3,
                                  bias=bias, gn_partial=part)
        fold = lambda: L.conv_gemm(a, w_up, batch=B, h_in=H, w_in=H, c_in=Cc, ksize=3, bias=bias, upsample2x=True, gn_partial=part)
        o1, o2 = big(), fold()
        torch.cuda.synchronize()
        err = ((o1.float() - o2.float()).norm() / o1.float().norm()).item()
        ok = err <= 5e-4
        bad += 0 if ok else 1
        line = f"up2x+conv3 B={B:2d} {H:3d}->{2 * H:3d} C={Cc:4d}: rel-L2(folded, materialised) {err:.2e} {'OK' if ok else 'MISMATCH'}"
        if not check_only:
            t1, t2 = timeit(big), timeit(fold)
            f9 = 2.0 * B * 4 * H * H * 9 * Cc * Cc
            line += f" | materialised {t1:8.1f} us ({f9 / t1 / 1e6:7.1f} TF/s) | folded {t2:8.1f} us ({f9 * 4 / 9 / t2 / 1e6:7.1f} TF/s executed) | x{t1 / t2:.2f}"
        print(line, flush=True)
    if not check_only:      # kernel choice on the short-K decoder upsamplers (K = 4 * C)
        for B, H, Cc in [(1, 256, 256), (8, 256, 256), (1, 128, 512), (1, 64, 512)]:
            a = torch.randn(B * H * H, Cc, device="cuda", generator=g).half()
            w_up = upsample_conv_weight((torch.randn(Cc, Cc, 3, 3, device="cuda", generator=g) / math.sqrt(9 * Cc)).half())
            bias = torch.randn(Cc, device="cuda", generator=g)
            part = torch.empty(L.gn_partial_numel(B, 4 * H * H), device="cuda")
            for name, kw in [("auto", {}), ("auto, no stats", dict(gn_partial=None)), ("single CTA 128x256", dict(cta_pair=1, tile_n=256, no_persistent=2)),
                             ("single CTA 128x128", dict(cta_pair=1, tile_n=128, no_persistent=2)), ("pair 256x128", dict(cta_pair=2, tile_n=128))]:
                args = dict(batch=B, h_in=H, w_in=H, c_in=Cc, ksize=3, bias=bias, upsample2x=True, gn_partial=part, split_k=1)
                args.update(kw)
                t = timeit(lambda: L.conv_gemm(a, w_up, **args))
                print(f"variant B={B} {H}->{2 * H} C={Cc} {name:20s}: {t:8.1f} us ({2.0 * B * 4 * H * H * 4 * Cc * Cc / t / 1e6:7.1f} TF/s executed)", flush=True)
    print("up_bench:", "ALL OK" if bad == 0 else f"{bad} MISMATCHES")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
