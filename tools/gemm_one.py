"""One conv/GEMM shape launched a few times, for `ncu --set full -k regex:conv_gemm -s 3 -c 1`.
usage: python tools/gemm_one.py conv3|lin B H Cin Cout [cta_pair [halo]]"""
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
else:
    a = torch.randn(H, Ci, device="cuda", generator=g).half()
    w = (torch.randn(Co, Ci, device="cuda", generator=g) / math.sqrt(Ci)).half()
    res = torch.randn(H, Co, device="cuda", generator=g).half()
    bias = torch.randn(Co, device="cuda", generator=g)
    f = lambda: L.conv_gemm(a, w, batch=1, h_in=1, w_in=H, c_in=Ci, bias=bias, residual=res, cta_pair=cp)
for _ in range(6):
    f()
torch.cuda.synchronize()
