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


def tma_store_bench():
    g = torch.Generator(device="cuda").manual_seed(1)
    for M, K, N, stats in [(262144, 64, 128, True), (1048576, 64, 128, True), (2097152, 64, 128, True), (16384, 320, 640, False), (16384, 320, 1280, False),
                           (65536, 320, 640, False), (131072, 320, 1280, False), (16384, 640, 640, False), (65536, 128, 128, False), (262144, 128, 256, False),
                           (65536, 1280, 1280, False)]:
        a = torch.randn(M, K, device="cuda", generator=g).half()
        w = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).half()
        bias = torch.randn(N, device="cuda", generator=g)
        part = torch.empty(L.gn_partial_numel(1, M), device="cuda") if stats else None
        out = torch.empty(M, N, device="cuda", dtype=torch.float16)
        args = dict(batch=1, h_in=1, w_in=M, c_in=K, bias=bias, out=out, gn_partial=part)
        t0 = timeit(lambda: L.conv_gemm(a, w, tma_store=1, **args))
        t1 = timeit(lambda: L.conv_gemm(a, w, tma_store=1, tile_n=128, no_persistent=2, split_k=1, cta_pair=1, **args))
        t2 = timeit(lambda: L.conv_gemm(a, w, tma_store=2, tile_n=128, no_persistent=2, split_k=1, cta_pair=1, **args))
        gb = 2.0 * (M * K + N * K + M * N) / 1e3
        print(f"gemm m{M}_k{K}_n{N}{' +stats' if stats else ''}: auto (row stores) {t0:7.1f} us | persistent 128 row stores {t1:7.1f} us ({gb / t1:6.0f} GB/s) | "
              f"TMA store {t2:7.1f} us ({gb / t2:6.0f} GB/s, {2.0 * M * K * N / t2 / 1e6:6.0f} TF/s)", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "tma":
        tma_store_bench()
        sys.exit(0)
    main()
