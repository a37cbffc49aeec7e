"""Device time of the small layout kernels of the step: concat + FreeU (one launch vs the two-pass kernels), image-in
(64 zero-padded channels + 3x3 conv vs 3x3 patches + K = 64 GEMM).   usage: python tools/misc_bench.py"""
import math
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from instantrestore_b200 import _lib as L
from instantrestore_b200.weights import conv_weight_khwc, patch_conv_weight


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
    g = torch.Generator(device="cuda").manual_seed(0)
    for B, H, Ch, Cs in [(1, 8, 1280, 1280), (4, 8, 1280, 1280), (8, 8, 1280, 1280), (1, 16, 1280, 1280), (4, 16, 1280, 1280),
                         (1, 16, 1280, 640), (4, 16, 1280, 640), (8, 16, 1280, 1280), (32, 16, 1280, 1280)]:
        hid = torch.randn(B * H * H, Ch, device="cuda", generator=g).half()
        sk = torch.randn(B * H * H, Cs, device="cuda", generator=g).half()
        out = torch.empty(B * H * H, Ch + Cs, device="cuda", dtype=torch.float16)
        f = lambda tp: L.concat_freeu(hid, sk, batch=B, h=H, w=H, backbone_scale=1.4, skip_scale=0.9, out=out, two_pass=tp)
        a = f(False).clone()
        same = torch.equal(a, f(True))
        t1, t2 = timeit(lambda: f(True)), timeit(lambda: f(False))
        print(f"concat+FreeU B={B:2d} {H:2d}x{H:2d} {Ch}+{Cs}: two-pass {t1:6.1f} us | one launch {t2:6.1f} us | bit-identical {same}", flush=True)
    for B in (1, 4, 8):
        img = (torch.rand(B, 3, 512, 512, device="cuda", generator=g) * 2 - 1).half()
        w = (torch.randn(128, 3, 3, 3, device="cuda", generator=g) / math.sqrt(27)).half()
        bias = torch.randn(128, device="cuda", generator=g)
        wk, wp = conv_weight_khwc(w, 64), patch_conv_weight(w)
        part = torch.empty(L.gn_partial_numel(B, 512 * 512), device="cuda")
        old = lambda: L.conv_gemm(L.image_in(img), wk, batch=B, h_in=512, w_in=512, c_in=64, ksize=3, bias=bias, gn_partial=part)
        new = lambda: L.conv_gemm(L.image_in_patches3x3(img), wp, batch=B, h_in=1, w_in=512 * 512, c_in=64, bias=bias, gn_partial=part)
        o1, o2 = old(), new()
        err = ((o1.float() - o2.float()).norm() / o1.float().norm()).item()
        t1, t2 = timeit(old), timeit(new)
        t3, t4 = timeit(lambda: L.image_in(img)), timeit(lambda: L.image_in_patches3x3(img))
        print(f"image-in + conv_in B={B}: padded 3x3 conv {t1:7.1f} us (layout {t3:6.1f}) | patches + K=64 GEMM {t2:7.1f} us (patches {t4:6.1f}) | rel-L2 {err:.2e}", flush=True)


if __name__ == "__main__":
    main()
