"""One conv/GEMM shape launched a few times, for `ncu --set full -k regex:conv_gemm -s 3 -c 1`.
usage: python tools/gemm_one.py conv3|lin|up|lin_ts B H Cin Cout [cta_pair [halo]]
  up: nearest-2x + 3x3 conv folded (upsample2x) with GroupNorm statistics; lin_ts: no residual, statistics, TMA-store epilogue (cta_pair = tma_store mode)"""
import math
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from instantrestore_b200 import _lib as L

kind, B, H, Ci, Co = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
cp = int(sys.argv[6]) if len(sys.argv) > 6 else 0
halo = int(sys.argv[7]) if len(sys.argv) > 7 else 0
g = torch.Generator(device="cuda").manual_seed(0)
if kind == "conv3":
    a = torch.randn(B * H * H, Ci, device="cuda", generator=g).half()
    w = (torch.randn(Co, 9 * Ci, device="cuda", generator=g) / math.sqrt(9 * Ci)).half()
    res = torch.randn(B * H * H, Co, device="cuda", generator=g).half()
    bias = torch.randn(Co, device="cuda", generator=g)
    f = lambda: L.conv_gemm(a, w, batch=B, h_in=H, w_in=H, c_in=Ci, ksize=3, bias=bias, residual=res, cta_pair=cp, halo=halo)
elif kind == "up":
    from instantrestore_b200.weights import upsample_conv_weight
    a = torch.randn(B * H * H, Ci, device="cuda", generator=g).half()
    w = upsample_conv_weight((torch.randn(Co, Ci, 3, 3, device="cuda", generator=g) / math.sqrt(9 * Ci)).half())
    bias = torch.randn(Co, device="cuda", generator=g)
    part = torch.empty(L.gn_partial_numel(B, 4 * H * H), device="cuda")
    f = lambda: L.conv_gemm(a, w, batch=B, h_in=H, w_in=H, c_in=Ci, ksize=3, bias=bias, upsample2x=True, gn_partial=part, cta_pair=cp)
elif kind == "lin_ts":
    a = torch.randn(H, Ci, device="cuda", generator=g).half()
    w = (torch.randn(Co, Ci, device="cuda", generator=g) / math.sqrt(Ci)).half()
    bias = torch.randn(Co, device="cuda", generator=g)
    part = torch.empty(L.gn_partial_numel(B, H // B), device="cuda")
    f = lambda: L.conv_gemm(a, w, batch=B, h_in=1, w_in=H // B, c_in=Ci, bias=bias, gn_partial=part, tma_store=cp)
else:
    a = torch.randn(H, Ci, device="cuda", generator=g).half()
    w = (torch.randn(Co, Ci, device="cuda", generator=g) / math.sqrt(Ci)).half()
    res = torch.randn(H, Co, device="cuda", generator=g).half()
    bias = torch.randn(Co, device="cuda", generator=g)
    f = lambda: L.conv_gemm(a, w, batch=1, h_in=1, w_in=H, c_in=Ci, bias=bias, residual=res, cta_pair=cp)
for _ in range(6):
    f()
torch.cuda.synchronize()
