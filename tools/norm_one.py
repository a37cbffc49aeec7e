"""One GroupNorm / LayerNorm shape launched a few times, for `ncu --set full -k regex:<kernel> -s 3 -c 1`.
usage: python tools/norm_one.py gn B HW C [fused: 0 auto | 1 three-kernel | 2 single-launch]   |   python tools/norm_one.py ln R C"""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from instantrestore_b200 import _lib as L

g = torch.Generator(device="cuda").manual_seed(0)
if sys.argv[1] == "gn":
    B, HW, C = (int(a) for a in sys.argv[2:5])
    fused = int(sys.argv[5]) if len(sys.argv) > 5 else 0
    x = torch.randn(B * HW, C, device="cuda", generator=g).half()
    gm, bt = torch.randn(C, device="cuda", generator=g), torch.randn(C, device="cuda", generator=g)
    out = torch.empty_like(x)
    f = lambda: L.groupnorm(x, gm, bt, batch=B, hw=HW, silu=True, out=out, fused=fused or None)
else:
    R, C = int(sys.argv[2]), int(sys.argv[3])
    x = torch.randn(R, C, device="cuda", generator=g).half()
    gm, bt = torch.randn(C, device="cuda", generator=g), torch.randn(C, device="cuda", generator=g)
    out = torch.empty_like(x)
    f = lambda: L.layernorm(x, gm, bt, out=out)
for _ in range(6):
    f()
torch.cuda.synchronize()
