"""For every ir_conv_gemm (op, shape) row of a bench.py kernel table: device time per launch of the automatic kernel choice
against the CTA-pair kernel forced at 256 x 128, 256 x 160 and 256 x 256 tiles (8 back-to-back launches in a CUDA graph, best
of 5 replays). Prints the rows where a forced choice wins, weighted by launches per step.
usage: python tools/pair_sweep.py profiles/r02af_kernel_table_b1.json [more tables]"""
import json
import math
import re
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from instantrestore_b200 import _lib as L


def timed(f):
    try:
        f()
    except RuntimeError:
        return None
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
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 8 * 1e3)
    return best


def main():
    g = torch.Generator(device="cuda").manual_seed(0)
    shapes = {}
    for path in sys.argv[1:]:
        for r in json.load(open(path))["rows"]:
            if r["op"] != "ir_conv_gemm":
                continue
            m_ = re.fullmatch(r"m(\d+)_k(\d+)_n(\d+)_ks(\d)s(\d)", r["shape"])
            if not m_:
                continue
            M, K, N, ks, st = map(int, m_.groups())
            if st != 1 or N % 128 and N % 160:
                continue
            key = (M, K, N, ks)
            shapes[key] = max(shapes.get(key, 0), r["launches_per_step"])
    total_gain = 0.0
    for (M, K, N, ks), n in sorted(shapes.items()):
        if ks == 3:
            ci = K // 9
            b = 1
            while int(math.isqrt(M // b)) ** 2 != M // b or (M // b) & (M // b - 1):
                b *= 2
                if b > 64:
                    break
            if b > 64:
                continue
            h = int(math.isqrt(M // b))
            a = torch.randn(M, ci, device="cuda", generator=g).half()
            w = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).half()
            args = dict(batch=b, h_in=h, w_in=h, c_in=ci, ksize=3)
        else:
            a = torch.randn(M, K, device="cuda", generator=g).half()
            w = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).half()
            args = dict(batch=1, h_in=1, w_in=M, c_in=K)
        bias = torch.randn(N, device="cuda", generator=g)
        res = torch.randn(M, N, device="cuda", generator=g).half()
        out = torch.empty(M, N, device="cuda", dtype=torch.float16)
        run = lambda **kw: timed(lambda: L.conv_gemm(a, w, bias=bias, residual=res, out=out, **args, **kw))
        t0 = run()
        alts = {tn: run(cta_pair=2, tile_n=tn, split_k=1) for tn in (128, 160, 256)}
        alts = {k: v for k, v in alts.items() if v is not None}
        if not alts or t0 is None:
            continue
        tn, tb = min(alts.items(), key=lambda kv: kv[1])
        flag = ""
        if tb < 0.95 * t0:
            total_gain += (t0 - tb) * n
            flag = f"  <-- pair {tn}: -{(t0 - tb) * n:6.1f} us/step"
        print(f"m{M}_k{K}_n{N}_ks{ks} x{n:3d}: auto {t0:7.2f} us | " + " | ".join(f"pair{k} {v:7.2f}" for k, v in alts.items()) + flag, flush=True)
    print(f"total gain of the marked rows: {total_gain:.1f} us per step")


if __name__ == "__main__":
    main()
