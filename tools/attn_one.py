"""One shared-attention shape launched a few times, for `ncu --set full -k regex:shared_attn -s 3 -c 1`.
usage: python tools/attn_one.py B H S own n_ref adain"""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from instantrestore_b200 import _lib as L

B, H, S, own, n_ref, adain = (int(a) for a in sys.argv[1:7])
g = torch.Generator(device="cuda").manual_seed(0)
C = H * 64
q = torch.randn(B * S, 3 * C, device="cuda", generator=g).half()
kw = {}
if own:
    kw.update(k_own=q[:, C:], v_own=q[:, 2 * C:], s_own=S)
if n_ref:
    kv = torch.randn(B * n_ref * S, 3 * C, device="cuda", generator=g).half()
    kw.update(k_ref=kv[:, C:], v_ref=kv[:, 2 * C:], n_ref=n_ref, s_ref=S)
    if adain:
        kw.update(adain_scale=torch.rand(B, n_ref, C, device="cuda") + 0.5, adain_shift=torch.randn(B, n_ref, C, device="cuda"))
out = torch.empty(B * S, C, device="cuda", dtype=torch.float16)
for _ in range(6):
    L.shared_attn(q, heads=H, scale=0.125, batch=B, s_q=S, out=out, **kw)
torch.cuda.synchronize()
