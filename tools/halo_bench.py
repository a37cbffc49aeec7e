"""A/B of the halo conv kernel (input slice staged once per tile, nine shifted views) against the tap-by-tap kernels.
usage: python tools/halo_bench.py [check]"""
import math
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from instantrestore_b200 import _lib as L

SHAPES = [  # (batch, H, Cin, Cout)
    (1, 128, 64, 128), (2, 512, 128, 128), (8, 512, 128, 128), (1, 128, 512, 512), (4, 128, 512, 512), (8, 128, 512, 512),
    (2, 256, 256, 256), (8, 256, 256, 256), (8, 256, 128, 256), (1, 256, 256, 128), (2, 128, 256, 512), (1, 512, 128, 128),
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
    for B, H, Ci, Co in SHAPES:
        a = torch.randn(B * H * H, Ci, device="cuda", generator=g).half()
        w = (torch.randn(Co, 9 * Ci, device="cuda", generator=g) / math.sqrt(9 * Ci)).half()
        bias = torch.randn(Co, device="cuda", generator=g)
        res = torch.randn(B * H * H, Co, device="cuda", generator=g).half()
        f = lambda h, cp=0: L.conv_gemm(a, w, batch=B, h_in=H, w_in=H, c_in=Ci, ksize=3, bias=bias, residual=res, halo=h, cta_pair=cp)
        flops = 2.0 * B * H * H * 9 * Ci * Co
        o1, o2 = f(1), f(2)
        torch.cuda.synchronize()
        err = ((o1.float() - o2.float()).norm() / o1.float().norm()).item()
        ok = err <= 3e-4
        bad += 0 if ok else 1
        line = f"conv3 B={B:2d} H={H:4d} {Ci:4d}->{Co:4d}: rel-L2(halo, taps) {err:.2e} {'OK' if ok else 'MISMATCH'}"
        if not check_only:
            t1, t2 = timeit(lambda: f(1)), timeit(lambda: f(2))
            line += f" | taps {t1:8.1f} us {flops / t1 / 1e6:7.1f} TF/s | halo {t2:8.1f} us {flops / t2 / 1e6:7.1f} TF/s"
            if Co % 256:      # 128-wide outputs: the CTA pair (default) against the single-CTA kernel
                o3 = f(2, 1)
                torch.cuda.synchronize()
                e3 = ((o3.float() - o2.float()).norm() / o2.float().norm()).item()
                t3 = timeit(lambda: f(2, 1))
                line += f" | single CTA {t3:8.1f} us {flops / t3 / 1e6:7.1f} TF/s (rel-L2 vs pair {e3:.1e})"
        print(line, flush=True)
    print("halo_bench:", "ALL OK" if bad == 0 else f"{bad} MISMATCHES")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
