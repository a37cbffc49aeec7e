"""GPU bring-up probe: runs every kernel of the C ABI against plain-torch fp32 math, one subprocess per case so a
trapped kernel cannot take the rest down. Writes gpurun_out/probe.jsonl. Usage: python tools/gpu_probe.py [case ...]"""
from __future__ import annotations

import json
import math
import os
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
OUT = ROOT / "gpurun_out"


def rel_err(a, b):
    import torch
    a = a.float(); b = b.float()
    return (torch.linalg.norm(a - b) / (torch.linalg.norm(b) + 1e-12)).item(), (a - b).abs().max().item()


def case_umma():
    import torch
    from instantrestore_b200 import _lib as L
    lib = L.load()
    res = []
    g = torch.Generator(device="cuda").manual_seed(0)
    a = (torch.randn(128, 64, device="cuda", generator=g)).half()
    for N in (64, 128, 256):
        # K-major B: [N, 64]
        b = torch.randn(N, 64, device="cuda", generator=g).half()
        ref = a.float() @ b.float().T
        for (lbo, sbo, kadv) in [(16, 1024, 32), (0, 1024, 32), (1024, 1024, 32)]:
            d = torch.full((128, N), float("nan"), device="cuda")
            rc = lib.ir_debug_umma(a.data_ptr(), b.data_ptr(), d.data_ptr(), N, 0, 16, 1024, lbo, sbo, 32, kadv, None)
            torch.cuda.synchronize()
            res.append(dict(kind="kmajor", N=N, lbo=lbo, sbo=sbo, kadv=kadv, rc=rc, err=rel_err(d, ref)))
        # MN-major B: [64 (k), N]
        bt = torch.randn(64, N, device="cuda", generator=g).half()
        ref = a.float() @ bt.float()
        for (lbo, sbo, kadv) in [(8192, 1024, 2048), (1024, 8192, 2048), (16, 1024, 2048), (8192, 1024, 256),
                                 (1024, 1024, 2048), (8192, 2048, 2048), (0, 1024, 2048)]:
            d = torch.full((128, N), float("nan"), device="cuda")
            rc = lib.ir_debug_umma(a.data_ptr(), bt.data_ptr(), d.data_ptr(), N, 1, 16, 1024, lbo, sbo, 32, kadv, None)
            torch.cuda.synchronize()
            res.append(dict(kind="mnmajor", N=N, lbo=lbo, sbo=sbo, kadv=kadv, rc=rc, err=rel_err(d, ref)))
    return res


def _gemm_ref(a2d, w, bias, residual, act):
    import torch
    y = a2d.float() @ w.float().T
    if bias is not None:
        y = y + bias
    y = y.half().float()
    if residual is not None:
        y = y + residual.float()
    return y


def case_gemm():
    import torch
    from instantrestore_b200 import _lib as L
    res = []
    g = torch.Generator(device="cuda").manual_seed(1)
    for (M, K, N, tile_n, use_bias, use_res) in [
        (4096, 320 + 0, 320, 0, True, True), (4096, 320, 320, 64, True, False), (77, 1024, 640, 0, False, False),
        (256, 1280, 1280, 0, True, True), (1024, 640, 640, 128, True, True), (1024, 640, 640, 256, True, True),
        (64, 1280, 1280, 0, True, False), (300, 64, 72, 0, True, True), (4096, 320, 4, 0, True, False),
        (1, 320, 1280, 0, True, False),
    ]:
        a = torch.randn(M, K, device="cuda", generator=g).half()
        w = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).half()
        bias = torch.randn(N, device="cuda", generator=g) if use_bias else None
        r = torch.randn(M, N, device="cuda", generator=g).half() if use_res else None
        try:
            t0 = time.time()
            out = L.conv_gemm(a, w, batch=1, h_in=1, w_in=M, c_in=K, bias=bias, residual=r, tile_n=tile_n)
            torch.cuda.synchronize()
            ref = _gemm_ref(a, w, bias, r, 0)
            res.append(dict(kind="linear", M=M, K=K, N=N, tile_n=tile_n, err=rel_err(out, ref), ms=(time.time() - t0) * 1e3))
        except Exception as e:  # noqa
            res.append(dict(kind="linear", M=M, K=K, N=N, tile_n=tile_n, error=str(e)))
            if "CUDA" in str(e) or "launch" in str(e):
                break
    # GEGLU
    for (M, K, N, tile_n) in [(1024, 640, 5120, 0), (256, 64, 256, 128), (4096, 320, 2560, 256)]:
        a = torch.randn(M, K, device="cuda", generator=g).half()
        w = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).half()
        bias = torch.randn(N, device="cuda", generator=g)
        h = (a.float() @ w.float().T + bias).half().float()
        val, gate = h[:, : N // 2], h[:, N // 2:]
        ref = val * torch.nn.functional.gelu(gate).half().float()
        # interleave rows: blocks of 64 value rows followed by their 64 gate rows
        idx = []
        for blk in range(N // 128):
            idx += list(range(blk * 64, blk * 64 + 64)) + list(range(N // 2 + blk * 64, N // 2 + blk * 64 + 64))
        idx = torch.tensor(idx, device="cuda")
        try:
            out = L.conv_gemm(a, w[idx].contiguous(), batch=1, h_in=1, w_in=M, c_in=K, bias=bias[idx].contiguous(),
                              act=L.IR_ACT_GEGLU, tile_n=tile_n)
            torch.cuda.synchronize()
            res.append(dict(kind="geglu", M=M, K=K, N=N, tile_n=tile_n, err=rel_err(out, ref)))
        except Exception as e:  # noqa
            res.append(dict(kind="geglu", M=M, K=K, N=N, error=str(e)))
    return res


def case_conv():
    import torch
    import torch.nn.functional as F
    from instantrestore_b200 import _lib as L
    res = []
    g = torch.Generator(device="cuda").manual_seed(2)
    for (B, H, W, Ci, Co, stride, tile_n) in [
        (1, 64, 64, 64, 64, 1, 0), (2, 16, 16, 128, 192, 1, 0), (3, 8, 8, 64, 128, 1, 0), (1, 4, 4, 64, 64, 1, 0),
        (5, 4, 4, 128, 64, 1, 0), (1, 64, 64, 320, 320, 1, 0), (1, 64, 64, 320, 320, 1, 64), (2, 32, 32, 640, 640, 1, 0),
        (1, 64, 64, 64, 64, 2, 0), (2, 32, 32, 128, 128, 2, 0), (3, 8, 8, 64, 64, 2, 0), (1, 16, 16, 1280, 1280, 1, 0),
        (1, 8, 8, 2560, 1280, 1, 0), (1, 128, 128, 64, 64, 1, 0), (1, 256, 256, 64, 128, 1, 0),
    ]:
        x = torch.randn(B, Ci, H, W, device="cuda", generator=g).half()
        w = (torch.randn(Co, Ci, 3, 3, device="cuda", generator=g) / math.sqrt(9 * Ci)).half()
        bias = torch.randn(Co, device="cuda", generator=g)
        ref = F.conv2d(x.float(), w.float(), bias, stride=stride, padding=1)
        ref = ref.permute(0, 2, 3, 1).reshape(-1, Co)
        a = x.permute(0, 2, 3, 1).contiguous().reshape(-1, Ci)
        wk = w.permute(0, 2, 3, 1).contiguous().reshape(Co, 9 * Ci)
        try:
            out = L.conv_gemm(a, wk, batch=B, h_in=H, w_in=W, c_in=Ci, ksize=3, stride=stride, bias=bias, tile_n=tile_n)
            torch.cuda.synchronize()
            res.append(dict(kind="conv3", B=B, H=H, W=W, Ci=Ci, Co=Co, stride=stride, tile_n=tile_n, err=rel_err(out, ref)))
        except Exception as e:  # noqa
            res.append(dict(kind="conv3", B=B, H=H, W=W, Ci=Ci, Co=Co, stride=stride, error=str(e)))
            if "CUDA" in str(e):
                break
    return res


def _attn_ref(q, k_chunks, v_chunks, heads, scale, kv_lens=None):
    """q [B,S,C]; chunks lists of [B,L,C] float; returns [B,S,C] fp32"""
    import torch
    B, S, Cc = q.shape
    k = torch.cat(k_chunks, 1); v = torch.cat(v_chunks, 1)
    qh = q.float().reshape(B, S, heads, 64).transpose(1, 2)
    kh = k.float().reshape(B, -1, heads, 64).transpose(1, 2)
    vh = v.float().reshape(B, -1, heads, 64).transpose(1, 2)
    p = torch.softmax(qh @ kh.transpose(-1, -2) * scale, -1)
    return (p @ vh).transpose(1, 2).reshape(B, S, Cc)


def case_attn():
    import torch
    from instantrestore_b200 import _lib as L
    res = []
    g = torch.Generator(device="cuda").manual_seed(3)
    scale = 0.125
    cfgs = [  # (B, H, S, own, n_ref, adain, s_own_override)
        (1, 2, 256, True, 0, False, None), (1, 1, 128, True, 0, False, None), (2, 2, 256, False, 2, False, None),
        (1, 2, 256, True, 2, False, None), (2, 3, 256, False, 3, True, None), (1, 2, 64, True, 1, True, None),
        (1, 2, 256, True, 0, False, 77), (1, 5, 1024, False, 4, True, None), (1, 5, 4096, False, 4, True, None),
        (2, 20, 256, True, 4, True, None),
    ]
    for (B, H, S, own, n_ref, adain, s_own_o) in cfgs:
        Cc = H * 64
        q = torch.randn(B, S, Cc, device="cuda", generator=g).half()
        s_own = s_own_o or S
        shared = s_own_o is not None
        kc, vc = [], []
        kw = {}
        if own:
            nb = 1 if shared else B
            ko = torch.randn(nb, s_own, Cc, device="cuda", generator=g).half()
            vo = torch.randn(nb, s_own, Cc, device="cuda", generator=g).half()
            kc.append(ko.expand(B, -1, -1)); vc.append(vo.expand(B, -1, -1).float())
            kw.update(k_own=ko.reshape(-1, Cc), v_own=vo.reshape(-1, Cc), s_own=s_own, own_shared=shared)
        if n_ref:
            kr = torch.randn(B, n_ref, S, Cc, device="cuda", generator=g).half()
            vr = (torch.randn(B, n_ref, S, Cc, device="cuda", generator=g) * 1.5 + 0.3).half()
            kw.update(k_ref=kr.reshape(-1, Cc), v_ref=vr.reshape(-1, Cc), n_ref=n_ref, s_ref=S)
            a_s = a_b = None
            if adain:
                a_s = (torch.rand(B, n_ref, Cc, device="cuda", generator=g) + 0.5).contiguous()
                a_b = torch.randn(B, n_ref, Cc, device="cuda", generator=g).contiguous()
                kw.update(adain_scale=a_s, adain_shift=a_b)
            for r in range(n_ref):
                kc.append(kr[:, r])
                vv = vr[:, r].float()
                if adain:
                    vv = vv * a_s[:, r, None, :] + a_b[:, r, None, :]
                vc.append(vv)
        ref = _attn_ref(q, kc, vc, H, scale)
        try:
            out = L.shared_attn(q.reshape(-1, Cc), heads=H, scale=scale, batch=B, s_q=S, **kw)
            torch.cuda.synchronize()
            res.append(dict(kind="attn", B=B, H=H, S=S, own=own, n_ref=n_ref, adain=adain, s_own=s_own,
                            err=rel_err(out.reshape(B, S, Cc), ref)))
        except Exception as e:  # noqa
            res.append(dict(kind="attn", B=B, H=H, S=S, own=own, n_ref=n_ref, error=str(e)))
            if "CUDA" in str(e):
                break
    return res


def case_misc():
    import torch
    import torch.nn.functional as F
    from instantrestore_b200 import _lib as L
    res = []
    g = torch.Generator(device="cuda").manual_seed(4)
    # groupnorm
    for (B, HW, Cc, silu) in [(2, 4096, 320, True), (1, 64, 2560, True), (3, 256, 64, False), (1, 1024, 960, True)]:
        x = (torch.randn(B, HW, Cc, device="cuda", generator=g) * 2 + 0.5).half()
        gm = torch.randn(Cc, device="cuda", generator=g); bt = torch.randn(Cc, device="cuda", generator=g)
        ref = F.group_norm(x.float().transpose(1, 2), 32, gm, bt, 1e-5).transpose(1, 2)
        if silu:
            ref = F.silu(ref)
        out = L.groupnorm(x.reshape(-1, Cc), gm, bt, batch=B, hw=HW, silu=silu)
        torch.cuda.synchronize()
        res.append(dict(kind="groupnorm", B=B, HW=HW, C=Cc, err=rel_err(out.reshape(B, HW, Cc), ref)))
    # layernorm
    for (R, Cc) in [(4096, 320), (77, 1280), (1000, 64)]:
        x = (torch.randn(R, Cc, device="cuda", generator=g) * 2 + 0.5).half()
        gm = torch.randn(Cc, device="cuda", generator=g); bt = torch.randn(Cc, device="cuda", generator=g)
        ref = F.layer_norm(x.float(), (Cc,), gm, bt, 1e-5)
        out = L.layernorm(x, gm, bt)
        torch.cuda.synchronize()
        res.append(dict(kind="layernorm", R=R, C=Cc, err=rel_err(out, ref)))
    # adain coeffs
    B, S, Cc, N = 2, 256, 128, 3
    vo = (torch.randn(B, S, Cc, device="cuda", generator=g) * 1.3 + 0.2).half()
    vr = (torch.randn(B, N, S, Cc, device="cuda", generator=g) * 0.7 - 0.4).half()
    vr[1, 2] = 0  # padded (zeroed) reference slot
    sc, sh = L.adain_coeffs(vo.reshape(-1, Cc), vr.reshape(-1, Cc), batch=B, s_own=S, n_ref=N, s_ref=S, channels=Cc)
    torch.cuda.synchronize()
    sm, ss = vo.float().mean(1, keepdim=True), vo.float().std(1, keepdim=True) + 1e-5
    cm, cs = vr.float().mean(2), vr.float().std(2) + 1e-5
    ref_sc = ss / cs; ref_sh = sm - cm * ref_sc
    res.append(dict(kind="adain_scale", err=rel_err(sc, ref_sc)))
    res.append(dict(kind="adain_shift", err=rel_err(sh, ref_sh)))
    # concat + freeu
    for (B, H, W, Ch, Cs, bs, ss_) in [(2, 8, 8, 128, 64, 1.4, 0.9), (1, 16, 16, 1280, 640, 1.6, 0.2), (2, 32, 32, 64, 32, 1.0, 1.0)]:
        hid = torch.randn(B, H * W, Ch, device="cuda", generator=g).half()
        sk = torch.randn(B, H * W, Cs, device="cuda", generator=g).half()
        out = L.concat_freeu(hid.reshape(-1, Ch), sk.reshape(-1, Cs), batch=B, h=H, w=W, backbone_scale=bs, skip_scale=ss_)
        torch.cuda.synchronize()
        hh = hid.float().clone(); hh[..., : Ch // 2] *= bs
        hh = hh.half().float()
        x = sk.float().reshape(B, H, W, Cs).permute(0, 3, 1, 2)
        if ss_ != 1.0:
            xf = torch.fft.fftshift(torch.fft.fftn(x, dim=(-2, -1)), dim=(-2, -1))
            mask = torch.ones_like(x)
            mask[..., H // 2 - 1: H // 2 + 1, W // 2 - 1: W // 2 + 1] = ss_
            x = torch.fft.ifftn(torch.fft.ifftshift(xf * mask, dim=(-2, -1)), dim=(-2, -1)).real
        ref = torch.cat([hh, x.permute(0, 2, 3, 1).reshape(B, H * W, Cs)], -1)
        res.append(dict(kind="concat_freeu", B=B, H=H, C=(Ch, Cs), err=rel_err(out.reshape(B, H * W, -1), ref)))
    # upsample
    x = torch.randn(2, 8, 8, 64, device="cuda", generator=g).half()
    out = L.upsample_nearest2x(x.reshape(-1, 64), batch=2, h=8, w=8)
    ref = F.interpolate(x.permute(0, 3, 1, 2).float(), scale_factor=2.0, mode="nearest").permute(0, 2, 3, 1)
    res.append(dict(kind="upsample", err=rel_err(out.reshape(2, 16, 16, 64), ref)))
    # latent in/out
    x = torch.randn(2, 4, 16, 16, device="cuda", generator=g); nz = torch.randn(2, 4, 16, 16, device="cuda", generator=g)
    out = L.latent_in(x, nz, 0.9, 0.3)
    ref = torch.zeros(2, 256, 64, device="cuda"); ref[..., :4] = (0.9 * x + 0.3 * nz).permute(0, 2, 3, 1).reshape(2, 256, 4)
    res.append(dict(kind="latent_in", err=rel_err(out.reshape(2, 256, 64), ref)))
    eps = torch.randn(2 * 256, 8, device="cuda", generator=g).half()
    out = L.latent_out(eps, x, 0.3, 1.1)
    ref = (x - 0.3 * eps[:, :4].float().reshape(2, 16, 16, 4).permute(0, 3, 1, 2)) * 1.1
    res.append(dict(kind="latent_out", err=rel_err(out, ref)))
    return res


CASES = {"umma": case_umma, "gemm": case_gemm, "conv": case_conv, "attn": case_attn, "misc": case_misc}


def main():
    OUT.mkdir(exist_ok=True)
    if len(sys.argv) >= 3 and sys.argv[1] == "--run":
        name = sys.argv[2]
        try:
            res = CASES[name]()
        except Exception as e:  # noqa
            import traceback
            res = [dict(kind=name, fatal=str(e), tb=traceback.format_exc()[-1500:])]
        with open(OUT / "probe.jsonl", "a") as f:
            for r in res:
                r["case"] = name
                f.write(json.dumps(r) + "\n")
                print(json.dumps(r))
        return
    names = sys.argv[1:] or list(CASES)
    for n in names:
        print(f"=== {n}", flush=True)
        try:
            subprocess.run([sys.executable, __file__, "--run", n], timeout=300)
        except subprocess.TimeoutExpired:
            with open(OUT / "probe.jsonl", "a") as f:
                f.write(json.dumps(dict(case=n, fatal="timeout")) + "\n")
            print("TIMEOUT", n)


if __name__ == "__main__":
    main()
