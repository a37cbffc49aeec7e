"""GPU parity, kernel level: every C-ABI entry point against plain fp32 torch math on the same fp16 inputs.
Tolerance (floating point path, north_star: 1e-3 relative fp16): relative L2 error <= 1e-3 for fp16 outputs
(one fp16 rounding of the result is 2.8e-4 RMS), 1e-5 for fp32 outputs."""
import math

import pytest
import torch
import torch.nn.functional as F

from conftest import rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-3


@pytest.fixture(scope="module")
def L():
    from instantrestore_b200 import _lib
    _lib.load()
    assert _lib.load().ir_check_device() == 0, _lib.load().ir_last_error_string()
    return _lib


def _gen(seed):
    return torch.Generator(device="cuda").manual_seed(seed)


# ------------------------------------------------------------------------------------------------ GEMM / conv
@pytest.mark.parametrize("M,K,N,tile_n,use_bias,use_res", [
    (4096, 320, 320, 0, True, True), (4096, 320, 320, 64, True, False), (77, 1024, 640, 0, False, False),
    (256, 1280, 1280, 0, True, True), (1024, 640, 640, 128, True, True), (1024, 640, 640, 256, True, True),
    (64, 1280, 1280, 0, True, False), (300, 64, 72, 0, True, True), (4096, 320, 4, 0, True, False),
    (1, 320, 1280, 0, True, False), (4096, 1280, 320, 0, True, True), (129, 64, 64, 0, False, False),
])
def test_linear(L, M, K, N, tile_n, use_bias, use_res):
    g = _gen(1)
    a = torch.randn(M, K, device="cuda", generator=g).half()
    w = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).half()
    bias = torch.randn(N, device="cuda", generator=g) if use_bias else None
    r = torch.randn(M, N, device="cuda", generator=g).half() if use_res else None
    out = L.conv_gemm(a, w, batch=1, h_in=1, w_in=M, c_in=K, bias=bias, residual=r, tile_n=tile_n)
    ref = a.float() @ w.float().T
    if bias is not None:
        ref = ref + bias
    if r is not None:
        ref = ref.half().float() + r.float()
    assert rel_l2(out, ref) <= TOL


@pytest.mark.parametrize("M,K,N,tile_n", [(1024, 640, 5120, 0), (256, 64, 256, 128), (4096, 320, 2560, 256), (77, 128, 1024, 0)])
def test_geglu_epilogue(L, M, K, N, tile_n):
    from instantrestore_b200.weights import geglu_interleave_index
    g = _gen(2)
    a = torch.randn(M, K, device="cuda", generator=g).half()
    w = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).half()
    bias = torch.randn(N, device="cuda", generator=g)
    h = (a.float() @ w.float().T + bias)
    ref = h[:, : N // 2] * F.gelu(h[:, N // 2:])
    idx = geglu_interleave_index(N).cuda()
    out = L.conv_gemm(a, w[idx].contiguous(), batch=1, h_in=1, w_in=M, c_in=K, bias=bias[idx].contiguous(),
                      act=L.IR_ACT_GEGLU, tile_n=tile_n)
    assert out.shape == (M, N // 2)
    assert rel_l2(out, ref) <= 2 * TOL     # two fp16 roundings (projection, gelu) before the product, as autocast does


@pytest.mark.parametrize("B,H,W,Ci,Co,stride,tile_n", [
    (1, 64, 64, 64, 64, 1, 0), (2, 16, 16, 128, 192, 1, 0), (3, 8, 8, 64, 128, 1, 0), (1, 4, 4, 64, 64, 1, 0),
    (5, 4, 4, 128, 64, 1, 0), (1, 64, 64, 320, 320, 1, 0), (1, 64, 64, 320, 320, 1, 64), (2, 32, 32, 640, 640, 1, 0),
    (1, 64, 64, 64, 64, 2, 0), (2, 32, 32, 128, 128, 2, 0), (3, 8, 8, 64, 64, 2, 0), (1, 16, 16, 1280, 1280, 1, 0),
    (1, 8, 8, 2560, 1280, 1, 0), (1, 64, 64, 320, 4, 1, 0), (3, 16, 16, 1920, 1280, 1, 0),
])
def test_conv3x3(L, B, H, W, Ci, Co, stride, tile_n):
    g = _gen(3)
    x = torch.randn(B, Ci, H, W, device="cuda", generator=g).half()
    w = (torch.randn(Co, Ci, 3, 3, device="cuda", generator=g) / math.sqrt(9 * Ci)).half()
    bias = torch.randn(Co, device="cuda", generator=g)
    ref = F.conv2d(x.float(), w.float(), bias, stride=stride, padding=1).permute(0, 2, 3, 1).reshape(-1, Co)
    a = x.permute(0, 2, 3, 1).contiguous().reshape(-1, Ci)
    wk = w.permute(0, 2, 3, 1).contiguous().reshape(Co, 9 * Ci)
    out = L.conv_gemm(a, wk, batch=B, h_in=H, w_in=W, c_in=Ci, ksize=3, stride=stride, bias=bias, tile_n=tile_n)
    assert rel_l2(out, ref) <= TOL


@pytest.mark.parametrize("M,K,N,split,tile_n", [
    (256, 1280, 1280, 4, 0), (256, 1280, 1280, 8, 0), (64, 1280, 1280, 8, 64), (1024, 640, 640, 2, 0),
    (256, 5120, 1280, 8, 128), (77, 1024, 320, 4, 64), (256, 1280, 1280, 0, 0), (64, 1280, 320, 0, 0),
])
def test_linear_split_k_cluster(L, M, K, N, split, tile_n):
    """Split-K across a thread-block cluster with the DSMEM reduce-scatter; split=0 exercises the auto heuristic."""
    g = _gen(21)
    a = torch.randn(M, K, device="cuda", generator=g).half()
    w = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).half()
    bias = torch.randn(N, device="cuda", generator=g)
    r = torch.randn(M, N, device="cuda", generator=g).half()
    out = L.conv_gemm(a, w, batch=1, h_in=1, w_in=M, c_in=K, bias=bias, residual=r, tile_n=tile_n, split_k=split)
    ref = (a.float() @ w.float().T + bias).half().float() + r.float()
    assert rel_l2(out, ref) <= TOL
    again = L.conv_gemm(a, w, batch=1, h_in=1, w_in=M, c_in=K, bias=bias, residual=r, tile_n=tile_n, split_k=split)
    assert torch.equal(out, again)                  # deterministic: fixed-order reduction, no atomics
    unsplit = L.conv_gemm(a, w, batch=1, h_in=1, w_in=M, c_in=K, bias=bias, residual=r, tile_n=tile_n, split_k=1)
    assert rel_l2(out, unsplit) <= 5e-4


@pytest.mark.parametrize("B,H,Ci,Co,stride,split", [
    (1, 8, 1280, 1280, 1, 8), (1, 16, 1280, 1280, 1, 8), (1, 8, 2560, 1280, 1, 0), (4, 8, 1280, 1280, 1, 0),
    (1, 16, 1280, 1280, 2, 4), (1, 32, 640, 640, 1, 2), (2, 4, 128, 64, 1, 2), (1, 32, 1280, 640, 1, 0),
])
def test_conv3x3_split_k_cluster(L, B, H, Ci, Co, stride, split):
    g = _gen(22)
    x = torch.randn(B, Ci, H, H, device="cuda", generator=g).half()
    w = (torch.randn(Co, Ci, 3, 3, device="cuda", generator=g) / math.sqrt(9 * Ci)).half()
    bias = torch.randn(Co, device="cuda", generator=g)
    ref = F.conv2d(x.float(), w.float(), bias, stride=stride, padding=1).permute(0, 2, 3, 1).reshape(-1, Co)
    a = x.permute(0, 2, 3, 1).contiguous().reshape(-1, Ci)
    wk = w.permute(0, 2, 3, 1).contiguous().reshape(Co, 9 * Ci)
    out = L.conv_gemm(a, wk, batch=B, h_in=H, w_in=H, c_in=Ci, ksize=3, stride=stride, bias=bias, split_k=split)
    assert rel_l2(out, ref) <= TOL


@pytest.mark.parametrize("M,K,N,use_res,geglu", [
    (4096, 512, 512, True, False), (640, 320, 256, True, False), (1000, 1280, 1280, False, False), (257, 64, 256, True, False),
    (20000, 640, 1024, True, False), (4096, 320, 2560, False, True), (777, 128, 512, False, True),
    (5000, 1152, 128, True, False), (900, 64, 640, False, False), (4096, 320, 320, True, False), (3000, 640, 960, True, False),
    (130, 64, 160, False, False),
])
def test_linear_cta_pair(L, M, K, N, use_res, geglu):
    """tcgen05 cta_group::2 kernel (256 x 256 tiles over two SMs): odd tile counts, ragged M, residual, GEGLU; must agree
    with fp32 math and, bit for bit, with the single-CTA kernel (same K order)."""
    from instantrestore_b200.weights import geglu_interleave_index
    g = _gen(31)
    a = torch.randn(M, K, device="cuda", generator=g).half()
    w = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).half()
    bias = torch.randn(N, device="cuda", generator=g)
    r = torch.randn(M, N, device="cuda", generator=g).half() if use_res else None
    h = a.float() @ w.float().T + bias
    if geglu:
        idx = geglu_interleave_index(N).cuda()
        kw = dict(batch=1, h_in=1, w_in=M, c_in=K, bias=bias[idx].contiguous(), act=L.IR_ACT_GEGLU)
        out = L.conv_gemm(a, w[idx].contiguous(), cta_pair=2, **kw)
        single = L.conv_gemm(a, w[idx].contiguous(), cta_pair=1, **kw)
        ref, tol = h[:, : N // 2] * F.gelu(h[:, N // 2:]), 2 * TOL
    else:
        kw = dict(batch=1, h_in=1, w_in=M, c_in=K, bias=bias, residual=r, split_k=1)
        out = L.conv_gemm(a, w, cta_pair=2, **kw)
        single = L.conv_gemm(a, w, cta_pair=1, **kw)
        ref, tol = (h.half().float() + r.float() if use_res else h), TOL
    assert rel_l2(out, ref) <= tol
    assert rel_l2(out, single) <= 2e-4


@pytest.mark.parametrize("B,H,Ci,Co,stride", [
    (1, 64, 512, 512, 1), (2, 32, 256, 256, 1), (3, 16, 1280, 1280, 1), (5, 8, 128, 256, 1), (1, 64, 128, 256, 2),
    (2, 128, 64, 256, 1), (2, 128, 64, 128, 1), (3, 32, 128, 128, 2), (2, 32, 320, 640, 1), (1, 64, 320, 320, 1),
    (4, 16, 640, 1280, 2),
])
def test_conv3x3_cta_pair(L, B, H, Ci, Co, stride):
    g = _gen(32)
    x = torch.randn(B, Ci, H, H, device="cuda", generator=g).half()
    w = (torch.randn(Co, Ci, 3, 3, device="cuda", generator=g) / math.sqrt(9 * Ci)).half()
    bias = torch.randn(Co, device="cuda", generator=g)
    ref = F.conv2d(x.float(), w.float(), bias, stride=stride, padding=1).permute(0, 2, 3, 1).reshape(-1, Co)
    a = x.permute(0, 2, 3, 1).contiguous().reshape(-1, Ci)
    wk = w.permute(0, 2, 3, 1).contiguous().reshape(Co, 9 * Ci)
    out = L.conv_gemm(a, wk, batch=B, h_in=H, w_in=H, c_in=Ci, ksize=3, stride=stride, bias=bias, cta_pair=2, split_k=1)
    assert rel_l2(out, ref) <= TOL
    again = L.conv_gemm(a, wk, batch=B, h_in=H, w_in=H, c_in=Ci, ksize=3, stride=stride, bias=bias, cta_pair=2, split_k=1)
    assert torch.equal(out, again)


@pytest.mark.parametrize("B,H,W,Ci,Co,use_res", [
    (1, 128, 128, 64, 128, True), (2, 128, 128, 128, 256, True), (1, 4, 256, 64, 128, False), (3, 2, 128, 128, 256, False),
    (1, 256, 256, 128, 128, True), (1, 6, 384, 64, 512, True), (5, 8, 128, 192, 384, True),
])
def test_conv3x3_halo(L, B, H, W, Ci, Co, use_res):
    """Halo kernel (input slice staged once, nine shifted views): image borders (TMA zero fill), several 128-pixel tiles
    per row, odd tile counts over the CTA pair, 128-wide (CTA pair by default, single CTA with cta_pair=1) and 256-wide
    (pair) N tiles."""
    g = _gen(33)
    x = torch.randn(B, Ci, H, W, device="cuda", generator=g).half()
    w = (torch.randn(Co, Ci, 3, 3, device="cuda", generator=g) / math.sqrt(9 * Ci)).half()
    bias = torch.randn(Co, device="cuda", generator=g)
    r = torch.randn(B * H * W, Co, device="cuda", generator=g).half() if use_res else None
    ref = F.conv2d(x.float(), w.float(), bias, padding=1).permute(0, 2, 3, 1).reshape(-1, Co)
    if use_res:
        ref = ref.half().float() + r.float()
    a = x.permute(0, 2, 3, 1).contiguous().reshape(-1, Ci)
    wk = w.permute(0, 2, 3, 1).contiguous().reshape(Co, 9 * Ci)
    kw = dict(batch=B, h_in=H, w_in=W, c_in=Ci, ksize=3, bias=bias, residual=r, split_k=1)
    out = L.conv_gemm(a, wk, halo=2, **kw)
    assert rel_l2(out, ref) <= TOL
    assert torch.equal(out, L.conv_gemm(a, wk, halo=2, **kw))
    if Co % 256:                                       # 128-wide N tiles: the single-CTA kernel computes the same sums in the same order
        single = L.conv_gemm(a, wk, halo=2, cta_pair=1, **kw)
        assert rel_l2(single, ref) <= TOL
        assert torch.equal(out, single)
    if (W & (W - 1)) == 0 and (H & (H - 1)) == 0:      # the tap-by-tap kernels need power-of-two images
        assert rel_l2(out, L.conv_gemm(a, wk, halo=1, **kw)) <= 3e-4      # same products, different fp32 summation order


def test_conv3x3_halo_pair128_strided_views(L):
    """The 128-wide halo pair moves residual and outputs through tensor maps: they may be column slices of wider buffers
    (row strides and column offsets that are multiples of 8 elements, the alignment ir_conv_gemm asks of every call)."""
    g = _gen(41)
    B, H, W, Ci, Co = 1, 128, 128, 128, 128
    a = torch.randn(B * H * W, Ci, device="cuda", generator=g).half()
    wk = (torch.randn(Co, 9 * Ci, device="cuda", generator=g) / math.sqrt(9 * Ci)).half()
    bias = torch.randn(Co, device="cuda", generator=g)
    res_wide = torch.randn(B * H * W, Co + 64, device="cuda", generator=g).half()
    res = res_wide[:, 8:8 + Co]                                  # stride Co + 64, 16-byte aligned column offset
    kw = dict(batch=B, h_in=H, w_in=W, c_in=Ci, ksize=3, bias=bias, split_k=1, halo=2)
    ref = L.conv_gemm(a, wk, residual=res.contiguous(), cta_pair=1, **kw)      # single CTA, contiguous
    out_wide = torch.full((B * H * W, 2 * Co), 7.0, device="cuda").half()
    L.conv_gemm(a, wk, residual=res, out=out_wide[:, Co:], **kw)               # pair: strided residual in, strided out
    assert torch.equal(out_wide[:, Co:], ref)
    assert bool((out_wide[:, :Co] == 7.0).all())                               # the neighbouring columns are untouched


@pytest.mark.parametrize("M,K,N,kw", [
    (4096, 64, 128, {}), (1000, 64, 128, {}), (16384, 320, 640, {}), (77, 128, 256, {}), (262144, 64, 128, dict(stats=True)),
    (8192, 320, 1280, dict(cols=True)), (4096, 640, 384, dict(act=True)), (128 * 149 + 5, 64, 128, {}),
])
def test_tma_store_epilogue_is_bit_identical(L, M, K, N, kw):
    """tma_store=2 (output tile packed into a swizzled shared-memory box, one cp.async.bulk.tensor store per half tile)
    against tma_store=1 (per-thread row stores) on the same persistent kernel family: bit-identical outputs and statistics,
    partial last M tiles clipped by the tensor map, nothing written past the tensor."""
    g = _gen(51)
    a = torch.randn(M, K, device="cuda", generator=g).half()
    w = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).half()
    bias = torch.randn(N, device="cuda", generator=g)
    args = dict(batch=1, h_in=1, w_in=M, c_in=K, bias=bias, tile_n=128, no_persistent=2, split_k=1, cta_pair=1)
    if kw.get("act"):
        args["act"] = L.IR_ACT_SILU
    outs, parts = [], []
    for mode in (1, 2):
        extra = {}
        if kw.get("stats"):
            extra = dict(gn_partial=torch.full((L.gn_partial_numel(1, M),), float("nan"), device="cuda"))
        if kw.get("cols"):
            extra = dict(col_partial=torch.full((M // 32, N - 2 * N // 4, 2), float("nan"), device="cuda"), col_begin=2 * N // 4)
        buf = torch.full((M + 64, N), 7.0, device="cuda", dtype=torch.float16)      # guard rows behind the tensor
        outs.append(L.conv_gemm(a, w, out=buf[:M], tma_store=mode, **args, **extra))
        parts.append(next(iter(extra.values())) if extra else None)
        assert (buf[M:] == 7.0).all()
    assert torch.equal(outs[0], outs[1])
    if parts[0] is not None:
        assert torch.isfinite(parts[1]).all() and torch.equal(parts[0], parts[1])
    ref = a.float() @ w.float().T + bias
    if kw.get("act"):
        ref = F.silu(ref.half().float())
    assert rel_l2(outs[1], ref) <= TOL


UP_CASES = [  # B, H, W, Ci, Co, kw
    (1, 8, 8, 64, 64, {}),                                  # one half-empty tile per phase, one tile per CTA
    (2, 8, 8, 128, 192, {}),                                # c_out not a multiple of the tile: the last N tile of a phase is clipped
    (3, 16, 8, 64, 128, {}),                                # non-square, odd tile count per phase
    (1, 16, 16, 1280, 1280, {}),                            # UNet up_blocks.1 upsampler at one identity (split-K clusters)
    (4, 8, 8, 1280, 1280, {}),                              # UNet up_blocks.0 upsampler of the four reference images
    (1, 32, 32, 640, 640, {}),                              # UNet up_blocks.2 upsampler
    (4, 32, 32, 640, 640, dict(cta_pair=2, split_k=1)),     # CTA-pair kernel, 160-wide tiles
    (1, 64, 64, 512, 512, {}),                              # VAE decoder up_blocks.0 upsampler (CTA pair, 256-wide tiles)
    (1, 128, 128, 256, 256, dict(cta_pair=2, split_k=1)),   # several 128-pixel tiles per image row
    (2, 64, 64, 128, 256, dict(cta_pair=1, no_persistent=2)),   # persistent single-CTA kernel
    (2, 32, 32, 128, 128, dict(cta_pair=1, no_persistent=1)),   # one tile per CTA, no split
    (1, 16, 16, 640, 128, dict(split_k=4)),                 # forced 4-way K split
    (5, 4, 4, 64, 64, {}),                                  # several images per tile, ragged last tile
]


@pytest.mark.parametrize("B,H,W,Ci,Co,kw", UP_CASES, ids=[f"b{c[0]}_{c[1]}x{c[2]}_{c[3]}to{c[4]}" + "".join(f"_{k}{v}" for k, v in c[5].items()) for c in UP_CASES])
def test_upsample2x_conv3x3(L, B, H, W, Ci, Co, kw):
    """ir_conv_gemm(upsample2x=1): nearest-2x + 3x3 conv (diffusers Upsample2D; reference block.py:2366,2476) as four
    2x2 sub-pixel convolutions on the low-resolution input, against F.interpolate + F.conv2d in fp32."""
    from instantrestore_b200.weights import upsample_conv_weight
    g = _gen(41)
    x = torch.randn(B, Ci, H, W, device="cuda", generator=g).half()
    w = (torch.randn(Co, Ci, 3, 3, device="cuda", generator=g) / math.sqrt(9 * Ci)).half()
    bias = torch.randn(Co, device="cuda", generator=g)
    up = F.interpolate(x.float(), scale_factor=2.0, mode="nearest")
    ref = F.conv2d(up, w.float(), bias, padding=1).permute(0, 2, 3, 1).reshape(-1, Co)
    a = x.permute(0, 2, 3, 1).contiguous().reshape(-1, Ci)
    w_up = upsample_conv_weight(w)
    assert w_up.shape == (4 * Co, 4 * Ci)
    out = L.conv_gemm(a, w_up, batch=B, h_in=H, w_in=W, c_in=Ci, ksize=3, bias=bias, upsample2x=True, **kw)
    assert out.shape == (4 * B * H * W, Co)
    assert rel_l2(out, ref) <= TOL
    assert torch.equal(out, L.conv_gemm(a, w_up, batch=B, h_in=H, w_in=W, c_in=Ci, ksize=3, bias=bias, upsample2x=True, **kw))
    # the materialised path (upsample kernel + 3x3 conv on the large tensor) computes the same products
    if (2 * H) & (2 * H - 1) == 0 and (2 * W) & (2 * W - 1) == 0:
        wk = w.permute(0, 2, 3, 1).contiguous().reshape(Co, 9 * Ci)
        big = L.conv_gemm(L.upsample_nearest2x(a, batch=B, h=H, w=W), wk, batch=B, h_in=2 * H, w_in=2 * W, c_in=Ci, ksize=3, bias=bias)
        assert rel_l2(out, big) <= 5e-4       # folded taps are rounded to fp16 once more


@pytest.mark.parametrize("B,H,W,Ci,Co,kw", [
    (2, 64, 64, 512, 512, {}),                              # CTA pair
    (1, 128, 128, 256, 256, {}),
    (2, 32, 32, 128, 256, dict(cta_pair=1, no_persistent=2)),   # persistent single-CTA kernel
    (1, 8, 8, 128, 128, {}),                                # one tile per CTA: statistics from the separate pass
])
def test_upsample2x_conv_emits_groupnorm_pass_a(L, B, H, W, Ci, Co, kw):
    """The folded upsampler's epilogue writes pass A of the next GroupNorm over the FULL-resolution image (slab slots are
    per phase: any partition of an image's pixels into 32-pixel slabs gives the same merged moments)."""
    from instantrestore_b200.weights import upsample_conv_weight
    g = _gen(42)
    a = torch.randn(B * H * W, Ci, device="cuda", generator=g).half()
    w = (torch.randn(Co, Ci, 3, 3, device="cuda", generator=g) / math.sqrt(9 * Ci)).half()
    bias = torch.randn(Co, device="cuda", generator=g)
    gamma, beta = torch.randn(Co, device="cuda", generator=g), torch.randn(Co, device="cuda", generator=g)
    hw = 4 * H * W
    part = torch.full((L.gn_partial_numel(B, hw),), float("nan"), device="cuda")
    args = dict(batch=B, h_in=H, w_in=W, c_in=Ci, ksize=3, bias=bias, upsample2x=True, **kw)
    w_up = upsample_conv_weight(w)
    y = L.conv_gemm(a, w_up, gn_partial=part, **args)
    assert torch.equal(y, L.conv_gemm(a, w_up, **args))
    assert torch.isfinite(part).all()
    fused = L.groupnorm(y, gamma, beta, batch=B, hw=hw, eps=1e-6, silu=True, partial_in=part, fused=1)
    plain = L.groupnorm(y, gamma, beta, batch=B, hw=hw, eps=1e-6, silu=True, fused=1)
    assert rel_l2(fused, plain) <= 2e-4
    x = y.float().view(B, hw, Co).permute(0, 2, 1)
    ref = F.silu(F.group_norm(x, 32, gamma, beta, eps=1e-6)).permute(0, 2, 1).reshape(-1, Co)
    assert rel_l2(fused, ref) <= TOL


def test_upsample2x_rejects_unsupported_combinations(L):
    a = torch.zeros(64, 64, device="cuda").half()
    w = torch.zeros(256, 256, device="cuda").half()
    r = torch.zeros(256, 64, device="cuda").half()
    p = L.ConvGemmParams(a=L.ptr(a), batch=1, h_in=8, w_in=8, c_in=64, a_row_stride=64, ksize=3, stride=1, w=L.ptr(w), c_out=64,
                         residual=L.ptr(r), res_row_stride=64, out=L.ptr(r), out_row_stride=64, upsample2x=1)
    import ctypes as C
    assert L.load().ir_conv_gemm(C.byref(p), None) != 0          # residual + upsample2x


@pytest.mark.parametrize("kind,B,H,W,Ci,Co,kw", [
    ("conv", 2, 128, 128, 128, 128, {}),                    # halo kernel, 128-wide CTA pair
    ("conv", 2, 128, 128, 128, 128, dict(cta_pair=1)),      # halo kernel, single CTA
    ("conv", 1, 128, 128, 256, 256, {}),                    # halo kernel, CTA pair
    ("conv", 4, 64, 64, 512, 512, dict(halo=1)),            # CTA-pair kernel (tap by tap)
    ("conv", 2, 64, 64, 128, 256, dict(halo=1, cta_pair=1)),  # persistent single-CTA kernel
    ("conv", 1, 16, 8, 128, 128, {}),                       # one tile per CTA: statistics from the separate pass
    ("conv", 2, 32, 32, 256, 512, dict(stride=2)),          # stride 2 (16 x 16 outputs)
    ("lin", 3, 1, 1024, 512, 512, {}),                      # 1 x 1 / linear with a residual, three "images" of 1024 tokens
])
def test_conv_emits_groupnorm_pass_a(L, kind, B, H, W, Ci, Co, kw):
    """ir_conv_gemm(gn_partial=...) + ir_groupnorm(partial_in=...) == the three-kernel GroupNorm of the same tensor."""
    g = _gen(34)
    stride = kw.pop("stride", 1)
    ks = 3 if kind == "conv" else 1
    a = torch.randn(B * H * W, Ci, device="cuda", generator=g).half()
    w = (torch.randn(Co, ks * ks * Ci, device="cuda", generator=g) / math.sqrt(ks * ks * Ci)).half()
    bias = torch.randn(Co, device="cuda", generator=g)
    hw = (H // stride) * (W // stride)
    r = torch.randn(B * hw, Co, device="cuda", generator=g).half()
    gamma, beta = torch.randn(Co, device="cuda", generator=g), torch.randn(Co, device="cuda", generator=g)
    assert L.gn_partial_supported(hw, Co)
    part = torch.full((L.gn_partial_numel(B, hw),), float("nan"), device="cuda")
    args = dict(batch=B, h_in=H, w_in=W, c_in=Ci, ksize=ks, stride=stride, bias=bias, residual=r, **kw)
    y = L.conv_gemm(a, w, gn_partial=part, **args)
    assert torch.equal(y, L.conv_gemm(a, w, **args))              # the statistics do not disturb the outputs
    assert torch.isfinite(part).all()                              # every (image, slab, group) slot was written
    fused = L.groupnorm(y, gamma, beta, batch=B, hw=hw, eps=1e-6, silu=True, partial_in=part, fused=1)   # fused=1: three-kernel path
    plain = L.groupnorm(y, gamma, beta, batch=B, hw=hw, eps=1e-6, silu=True, fused=1)
    assert rel_l2(fused, plain) <= 2e-4
    x = y.float().view(B, hw, Co).permute(0, 2, 1)
    ref = F.silu(F.group_norm(x, 32, gamma, beta, eps=1e-6)).permute(0, 2, 1).reshape(-1, Co)
    assert rel_l2(fused, ref) <= TOL
    # slab moments against fp32 math on the stored outputs
    pm = part.view(B, hw // 32, 32, 2)
    yv = y.float().view(B, hw // 32, 32, 32, Co // 32)               # [image, slab, pixel, group, channel in group]
    mean = yv.mean(dim=(2, 4))
    m2 = ((yv - mean[:, :, None, :, None]) ** 2).sum(dim=(2, 4))
    assert float((pm[..., 0] - mean).abs().max()) <= 1e-4
    assert rel_l2(pm[..., 1], m2) <= 1e-3


@pytest.mark.parametrize("kind,B,H,Ci,Co", [("conv", 2, 128, 128, 128), ("conv", 1, 128, 256, 256), ("conv", 4, 64, 320, 320),
                                           ("conv", 2, 32, 640, 640), ("lin", 1, 4096, 320, 320), ("lin", 1, 1000, 512, 512)])
def test_epilogue_256bit_io_is_bit_identical(L, kind, B, H, Ci, Co):
    """256-bit residual loads / output stores (auto when rows are 32-byte aligned) against the 128-bit path, and a
    misaligned output view (column offset of 8 channels) that has to fall back."""
    g = _gen(35)
    ks = 3 if kind == "conv" else 1
    M = B * H * H if kind == "conv" else H
    a = torch.randn(M, Ci, device="cuda", generator=g).half()
    w = (torch.randn(Co, ks * ks * Ci, device="cuda", generator=g) / math.sqrt(ks * ks * Ci)).half()
    bias = torch.randn(Co, device="cuda", generator=g)
    r = torch.randn(M, Co, device="cuda", generator=g).half()
    args = dict(batch=B, h_in=H, w_in=H, c_in=Ci, ksize=3, bias=bias, residual=r) if kind == "conv" else \
        dict(batch=1, h_in=1, w_in=M, c_in=Ci, bias=bias, residual=r)
    wide, narrow = L.conv_gemm(a, w, wide_io=0, **args), L.conv_gemm(a, w, wide_io=1, **args)
    assert torch.equal(wide, narrow)
    buf = torch.zeros(M, Co + 16, device="cuda", dtype=torch.float16)
    view = buf[:, 8:8 + Co]                                # rows start 16 bytes off a 32-byte boundary
    L.conv_gemm(a, w, out=view, **args)
    assert torch.equal(view, narrow) and float(buf[:, :8].abs().max()) == 0 and float(buf[:, 8 + Co:].abs().max()) == 0


def test_conv_rejects_bad_shapes(L):
    a = torch.zeros(64, 60, device="cuda", dtype=torch.float16)
    w = torch.zeros(64, 60, device="cuda", dtype=torch.float16)
    with pytest.raises(RuntimeError, match="c_in"):
        L.conv_gemm(a, w, batch=1, h_in=1, w_in=64, c_in=60)


# ------------------------------------------------------------------------------------------------ attention
def _attn_ref(q, k_chunks, v_chunks, heads, scale):
    B, S, C = q.shape
    k, v = torch.cat(k_chunks, 1), torch.cat(v_chunks, 1)
    qh = q.float().reshape(B, S, heads, 64).transpose(1, 2)
    kh = k.float().reshape(B, -1, heads, 64).transpose(1, 2)
    vh = v.float().reshape(B, -1, heads, 64).transpose(1, 2)
    p = torch.softmax(qh @ kh.transpose(-1, -2) * scale, -1)
    return (p @ vh).transpose(1, 2).reshape(B, S, C), p


ATTN_CFGS = [  # (B, H, S, own, n_ref, adain, s_own_override)
    (1, 2, 256, True, 0, False, None), (1, 1, 128, True, 0, False, None), (2, 2, 256, False, 2, False, None),
    (1, 2, 256, True, 2, False, None), (2, 3, 256, False, 3, True, None), (1, 2, 64, True, 1, True, None),
    (1, 2, 256, True, 0, False, 77), (1, 5, 1024, False, 4, True, None), (1, 5, 4096, False, 4, True, None),
    (2, 20, 256, True, 4, True, None), (1, 10, 1024, False, 1, False, None), (1, 5, 1024, True, 8, True, None),
    (3, 2, 64, False, 2, True, None),
]


def _attn_case(L, B, H, S, own, n_ref, adain, s_own_o, seed=4, want_mass=False):
    g = _gen(seed)
    C = H * 64
    q = torch.randn(B, S, C, device="cuda", generator=g).half()
    s_own = s_own_o or S
    shared = s_own_o is not None
    kc, vc, kw = [], [], {}
    if own:
        nb = 1 if shared else B
        ko = torch.randn(nb, s_own, C, device="cuda", generator=g).half()
        vo = torch.randn(nb, s_own, C, device="cuda", generator=g).half()
        kc.append(ko.expand(B, -1, -1)); vc.append(vo.expand(B, -1, -1).float())
        kw.update(k_own=ko.reshape(-1, C), v_own=vo.reshape(-1, C), s_own=s_own, own_shared=shared)
    if n_ref:
        kr = torch.randn(B, n_ref, S, C, device="cuda", generator=g).half()
        vr = (torch.randn(B, n_ref, S, C, device="cuda", generator=g) * 1.5 + 0.3).half()
        kw.update(k_ref=kr.reshape(-1, C), v_ref=vr.reshape(-1, C), n_ref=n_ref, s_ref=S)
        if adain:
            a_s = (torch.rand(B, n_ref, C, device="cuda", generator=g) + 0.5).contiguous()
            a_b = torch.randn(B, n_ref, C, device="cuda", generator=g).contiguous()
            kw.update(adain_scale=a_s, adain_shift=a_b)
        for r in range(n_ref):
            kc.append(kr[:, r])
            vv = vr[:, r].float()
            if adain:
                vv = vv * a_s[:, r, None, :] + a_b[:, r, None, :]
            vc.append(vv)
    ref, p = _attn_ref(q, kc, vc, H, 0.125)
    return q, kw, ref, p


@pytest.mark.parametrize("cfg", ATTN_CFGS, ids=[str(c) for c in ATTN_CFGS])
def test_shared_attention(L, cfg):
    B, H, S, own, n_ref, adain, s_own_o = cfg
    q, kw, ref, _ = _attn_case(L, *cfg)
    out = L.shared_attn(q.reshape(-1, H * 64), heads=H, scale=0.125, batch=B, s_q=S, **kw)
    assert rel_l2(out.reshape(B, S, -1), ref) <= TOL


def test_shared_attention_reads_fused_qkv_in_place(L):
    """q/k/v as column slices of one [B*S, 3C] projection output (what the engine passes): no copies."""
    g = _gen(5)
    B, H, S = 2, 2, 256
    C = H * 64
    qkv = torch.randn(B * S, 3 * C, device="cuda", generator=g).half()
    out = L.shared_attn(qkv, heads=H, scale=0.125, batch=B, s_q=S, k_own=qkv[:, C:], v_own=qkv[:, 2 * C:], s_own=S)
    q, k, v = (qkv[:, i * C:(i + 1) * C].reshape(B, S, C) for i in range(3))
    ref, _ = _attn_ref(q, [k], [v.float()], H, 0.125)
    assert rel_l2(out.reshape(B, S, C), ref) <= TOL


def test_attention_linearity_in_values(L):
    """Size-independent property at a full-size layer (S=4096, 5 heads, 4 refs): attention is linear in V."""
    g = _gen(6)
    B, H, S, N = 1, 5, 4096, 4
    C = H * 64
    q = torch.randn(B * S, C, device="cuda", generator=g).half()
    k = torch.randn(B * N * S, C, device="cuda", generator=g).half()
    v1 = torch.randn(B * N * S, C, device="cuda", generator=g).half()
    v2 = torch.randn(B * N * S, C, device="cuda", generator=g).half()
    f = lambda v: L.shared_attn(q, heads=H, scale=0.125, batch=B, s_q=S, k_ref=k, v_ref=v, n_ref=N, s_ref=S).float()
    vsum = (v1.float() + v2.float()).half()
    lhs, rhs = f(vsum), f(v1) + f(v2)
    assert rel_l2(lhs, rhs) <= 3e-3     # three independently rounded fp16 results + rounded v1+v2


def test_attention_constant_values_give_constant(L):
    """softmax rows sum to one: V == c  =>  O == c, at the largest layer with 8 references."""
    B, H, S, N = 1, 5, 4096, 8
    C = H * 64
    g = _gen(7)
    q = torch.randn(B * S, C, device="cuda", generator=g).half()
    k = torch.randn(B * N * S, C, device="cuda", generator=g).half()
    v = torch.full((B * N * S, C), 0.75, device="cuda", dtype=torch.float16)
    out = L.shared_attn(q, heads=H, scale=0.125, batch=B, s_q=S, k_ref=k, v_ref=v, n_ref=N, s_ref=S)
    assert float((out.float() - 0.75).abs().max()) <= 2e-3


# ------------------------------------------------------------------------------------------------ norms & helpers
@pytest.mark.parametrize("B,HW,C,silu", [(2, 4096, 320, True), (1, 64, 2560, True), (3, 256, 64, False), (1, 1024, 960, True), (2, 256, 1920, True)])
def test_groupnorm(L, B, HW, C, silu):
    g = _gen(8)
    x = (torch.randn(B, HW, C, device="cuda", generator=g) * 2 + 0.5).half()
    gm, bt = torch.randn(C, device="cuda", generator=g), torch.randn(C, device="cuda", generator=g)
    ref = F.group_norm(x.float().transpose(1, 2), 32, gm, bt, 1e-5).transpose(1, 2)
    if silu:
        ref = F.silu(ref)
    out = L.groupnorm(x.reshape(-1, C), gm, bt, batch=B, hw=HW, silu=silu)
    assert rel_l2(out.reshape(B, HW, C), ref) <= TOL


# GroupNorm shapes with a single-launch plan (tensors up to 12 MB): UNet (B = 1 main pass, B = 4 reference pass; concat
# widths 960 / 1920 / 2560 of the up blocks), VAE 64 x 64 / 128 x 128 levels (512 / 256 channels; 16 and 32 vectors per
# thread), plus odd sizes
@pytest.mark.parametrize("B,HW,C,silu", [
    (1, 4096, 320, True), (4, 4096, 320, True), (1, 4096, 640, True), (1, 4096, 960, True), (2, 1024, 640, False),
    (1, 1024, 1280, True), (1, 1024, 1920, True), (4, 256, 1280, True), (1, 256, 2560, True), (1, 64, 1280, True),
    (1, 64, 2560, True), (2, 4096, 512, True), (1, 8192, 512, True), (1, 16384, 256, True), (9, 4096, 160, True),
    (3, 256, 64, False), (2, 48, 320, True), (1, 16, 128, True)])
def test_groupnorm_single_launch(L, B, HW, C, silu):
    """gn_fused_kernel (cluster + DSMEM merge, tile in registers) vs fp32 torch AND vs the three-kernel path; strided
    input / output views; deterministic."""
    assert L.gn_fused_supported(B, HW, C)
    g = _gen(81)
    xs = (torch.randn(B * HW, C + 64, device="cuda", generator=g) * 2 + 0.5).half()
    x = xs[:, 32:32 + C]                                           # row stride C + 64, 64-byte column offset
    gm, bt = torch.randn(C, device="cuda", generator=g), torch.randn(C, device="cuda", generator=g)
    ref = F.group_norm(x.float().view(B, HW, C).transpose(1, 2), 32, gm, bt, 1e-5).transpose(1, 2)
    if silu:
        ref = F.silu(ref)
    outs = torch.full((B * HW, C + 8), float("nan"), device="cuda", dtype=torch.float16)
    out = L.groupnorm(x, gm, bt, batch=B, hw=HW, silu=silu, fused=2, out=outs[:, :C])
    assert torch.isfinite(out).all()
    assert rel_l2(out.reshape(B, HW, C), ref) <= TOL
    three = L.groupnorm(x, gm, bt, batch=B, hw=HW, silu=silu, fused=1)
    assert rel_l2(out, three) <= 3e-4                              # same math, different summation trees
    again = L.groupnorm(x, gm, bt, batch=B, hw=HW, silu=silu, fused=2)
    assert torch.equal(again, out.contiguous())
    # in place (out aliases x) is allowed: every thread holds its tile in registers before it stores
    xc = x.contiguous()
    inplace = L.groupnorm(xc, gm, bt, batch=B, hw=HW, silu=silu, fused=2, out=xc)
    assert torch.equal(inplace, out.contiguous())


def test_groupnorm_large_tensors_keep_the_three_kernel_path(L):
    assert not L.gn_fused_supported(1, 512 * 512, 128)
    assert not L.gn_fused_supported(4, 256 * 256, 256)
    assert not L.gn_fused_supported(32, 4096, 320)          # 84 MB: bandwidth-bound, the streaming path is faster
    with pytest.raises(RuntimeError):
        x = torch.zeros(256 * 256, 256, device="cuda", dtype=torch.float16)
        L.groupnorm(x, torch.ones(256, device="cuda"), torch.zeros(256, device="cuda"), batch=1, hw=256 * 256, fused=2)


@pytest.mark.parametrize("R,C", [(4096, 320), (77, 1280), (1000, 64), (1, 640)])
def test_layernorm(L, R, C):
    g = _gen(9)
    x = (torch.randn(R, C, device="cuda", generator=g) * 2 + 0.5).half()
    gm, bt = torch.randn(C, device="cuda", generator=g), torch.randn(C, device="cuda", generator=g)
    out = L.layernorm(x, gm, bt)
    assert rel_l2(out, F.layer_norm(x.float(), (C,), gm, bt, 1e-5)) <= TOL


def test_adain_coeffs_including_zeroed_slot(L):
    g = _gen(10)
    B, S, C, N = 2, 256, 128, 3
    vo = (torch.randn(B, S, C, device="cuda", generator=g) * 1.3 + 0.2).half()
    vr = (torch.randn(B, N, S, C, device="cuda", generator=g) * 0.7 - 0.4).half()
    vr[1, 2] = 0       # padded slot: std 0 -> scale = style_std / eps, shift = style_mean (reference quirk)
    sc, sh = L.adain_coeffs(vo.reshape(-1, C), vr.reshape(-1, C), batch=B, s_own=S, n_ref=N, s_ref=S, channels=C)
    sm, ss = vo.float().mean(1, keepdim=True), vo.float().std(1, keepdim=True) + 1e-5
    cm, cs = vr.float().mean(2), vr.float().std(2) + 1e-5
    ref_sc = ss / cs
    assert rel_l2(sc, ref_sc) <= 1e-5
    assert rel_l2(sh, sm - cm * ref_sc) <= 1e-5
    assert torch.allclose(sh[1, 2], sm[1, 0], rtol=0, atol=1e-6)


# the QKV projection shapes of the nine shared layers (C = 320 / 640 / 1280; B = 1 main pass, B*N = 4 reference pass, a
# B = 8 step) and forced kernel variants, so that every epilogue that can emit the column moments is exercised
@pytest.mark.parametrize("M,C,kw", [
    (4096, 320, {}), (1024, 640, {}), (256, 1280, {}),                       # one tile per CTA / split-K cluster (B = 1)
    (16384, 320, {}), (4096, 640, {}), (1024, 1280, {}),                     # persistent 160- / 256-wide tiles (B*N = 4)
    (32768, 320, {}), (8192, 1280, {}),                                      # B = 8 step
    (4096, 640, dict(cta_pair=2)), (2048, 320, dict(cta_pair=2)),            # CTA-pair kernel, 256- and 160-wide tiles
    (512, 640, dict(split_k=2)), (256, 1280, dict(split_k=4, tile_n=64)),     # split-K epilogue (8-column statistics)
    (1024, 320, dict(no_persistent=2, tile_n=64)), (2048, 640, dict(no_persistent=2, tile_n=128, m_sub=0)),
])
def test_gemm_epilogue_emits_adain_column_moments(L, M, C, kw):
    """ir_conv_gemm(col_partial=...): (mean, M2) per (32-token slab, V column) of the stored outputs, and ir_adain_coeffs
    from those moments == ir_adain_coeffs from a pass over V == fp32 torch (unbiased std, eps on the std)."""
    g = _gen(41)
    x = torch.randn(M, C, device="cuda", generator=g).half()
    w = (torch.randn(3 * C, C, device="cuda", generator=g) / math.sqrt(C)).half()
    part = torch.full((M // 32, C, 2), float("nan"), device="cuda")
    assert L.col_partial_supported(M, 3 * C, 2 * C)
    args = dict(batch=1, h_in=1, w_in=M, c_in=C, **kw)
    qkv = L.conv_gemm(x, w, col_partial=part, col_begin=2 * C, **args)
    assert torch.equal(qkv, L.conv_gemm(x, w, **args))                         # the statistics do not disturb the outputs
    assert torch.isfinite(part).all()                                          # every (slab, column) slot was written
    v = qkv[:, 2 * C:].float().view(M // 32, 32, C)
    mean = v.mean(1)
    m2 = ((v - mean[:, None]) ** 2).sum(1)
    assert float((part[..., 0] - mean).abs().max()) <= 1e-5
    assert rel_l2(part[..., 1], m2) <= 1e-4
    again = torch.empty_like(part)
    L.conv_gemm(x, w, col_partial=again, col_begin=2 * C, **args)
    assert torch.equal(again, part)                                            # fixed reduction tree: deterministic


@pytest.mark.parametrize("B,N,S,C", [(1, 4, 4096, 320), (2, 3, 1024, 640), (1, 8, 256, 1280)])
def test_adain_coeffs_from_epilogue_moments(L, B, N, S, C):
    g = _gen(42)
    xo = (torch.randn(B * S, C, device="cuda", generator=g) * 1.5 + 0.3).half()
    xr = (torch.randn(B * N * S, C, device="cuda", generator=g) * 0.7 - 0.2).half()
    w = (torch.randn(3 * C, C, device="cuda", generator=g) / math.sqrt(C)).half()
    po = torch.empty((B * S // 32, C, 2), device="cuda")
    pr = torch.empty((B * N * S // 32, C, 2), device="cuda")
    qo = L.conv_gemm(xo, w, batch=1, h_in=1, w_in=B * S, c_in=C, col_partial=po, col_begin=2 * C)
    qr = L.conv_gemm(xr, w, batch=1, h_in=1, w_in=B * N * S, c_in=C, col_partial=pr, col_begin=2 * C)
    # pad the last reference slot of the last identity exactly like the pipeline does: V rows and their moments zeroed
    qr.view(B, N, S, 3 * C)[B - 1, N - 1, :, C:].zero_()
    pr.view(B, N, -1)[B - 1, N - 1].zero_()
    sc, sh = L.adain_coeffs(None, None, batch=B, s_own=S, n_ref=N, s_ref=S, channels=C, own_partial=po, ref_partial=pr)
    sc2, sh2 = L.adain_coeffs(qo[:, 2 * C:], qr[:, 2 * C:], batch=B, s_own=S, n_ref=N, s_ref=S, channels=C)
    assert rel_l2(sc, sc2) <= 2e-5 and rel_l2(sh, sh2) <= 2e-5
    vo = qo[:, 2 * C:].float().view(B, S, C)
    vr = qr[:, 2 * C:].float().view(B, N, S, C)
    sm, ss = vo.mean(1), vo.std(1) + 1e-5
    cm, cs = vr.mean(2), vr.std(2) + 1e-5
    ref_sc = ss[:, None] / cs
    assert rel_l2(sc, ref_sc) <= 2e-5
    assert rel_l2(sh, sm[:, None] - cm * ref_sc) <= 2e-5


@pytest.mark.parametrize("B,H,Ch,Cs,bs,ss", [(2, 8, 128, 64, 1.4, 0.9), (1, 16, 1280, 640, 1.6, 0.2), (2, 32, 64, 32, 1.0, 1.0), (1, 8, 1280, 1280, 1.4, 0.9)])
def test_concat_freeu(L, B, H, Ch, Cs, bs, ss):
    g = _gen(11)
    W = H
    hid = torch.randn(B, H * W, Ch, device="cuda", generator=g).half()
    sk = torch.randn(B, H * W, Cs, device="cuda", generator=g).half()
    out = L.concat_freeu(hid.reshape(-1, Ch), sk.reshape(-1, Cs), batch=B, h=H, w=W, backbone_scale=bs, skip_scale=ss)
    hh = hid.float().clone()
    hh[..., : Ch // 2] *= bs
    x = sk.float().reshape(B, H, W, Cs).permute(0, 3, 1, 2)
    if ss != 1.0:
        xf = torch.fft.fftshift(torch.fft.fftn(x, dim=(-2, -1)), dim=(-2, -1))
        mask = torch.ones_like(x)
        mask[..., H // 2 - 1: H // 2 + 1, W // 2 - 1: W // 2 + 1] = ss
        x = torch.fft.ifftn(torch.fft.ifftshift(xf * mask, dim=(-2, -1)), dim=(-2, -1)).real
    ref = torch.cat([hh, x.permute(0, 2, 3, 1).reshape(B, H * W, Cs)], -1)
    assert rel_l2(out.reshape(B, H * W, -1), ref) <= TOL
    # the one-launch kernel of the 8 x 8 / 16 x 16 stages performs the two-pass kernels' operations in the same order
    two = L.concat_freeu(hid.reshape(-1, Ch), sk.reshape(-1, Cs), batch=B, h=H, w=W, backbone_scale=bs, skip_scale=ss, two_pass=True)
    assert rel_l2(two.reshape(B, H * W, -1), ref) <= TOL
    assert float((out.float() - two.float()).abs().max()) <= 4e-3      # same sums; fp32 contraction may differ by an fp16 ulp


@pytest.mark.parametrize("B,C,H,W,dt", [(1, 3, 16, 16, torch.float16), (2, 3, 8, 24, torch.float32), (1, 3, 128, 128, torch.float16), (3, 7, 4, 4, torch.float32),
                                        (2, 3, 64, 32, torch.float32), (1, 3, 5, 7, torch.float16), (1, 3, 512, 512, torch.float16)])
def test_image_patches_conv_in(L, B, C, H, W, dt):
    """ir_image_in_patches3x3 + a K = 64 GEMM == the VAE encoder's conv_in (3x3, padding 1) on the NCHW image: patches are
    exact copies (bit-exact against F.unfold), the GEMM within the kernel tolerance of F.conv2d."""
    from instantrestore_b200.weights import patch_conv_weight
    g = _gen(45)
    img = (torch.rand(B, C, H, W, device="cuda", generator=g) * 2 - 1).to(dt)
    patches = L.image_in_patches3x3(img)
    assert patches.shape == (B * H * W, 64)
    un = F.unfold(img.float(), 3, padding=1).view(B, C, 9, H * W).permute(0, 3, 2, 1).reshape(B * H * W, 9 * C)     # (tap, ch) order
    assert torch.equal(patches[:, : 9 * C], un.half())
    assert not patches[:, 9 * C:].any()
    w = (torch.randn(128, C, 3, 3, device="cuda", generator=g) / math.sqrt(9 * C)).half()
    bias = torch.randn(128, device="cuda", generator=g)
    out = L.conv_gemm(patches, patch_conv_weight(w), batch=1, h_in=1, w_in=B * H * W, c_in=64, bias=bias)
    ref = F.conv2d(img.half().float(), w.float(), bias, padding=1).permute(0, 2, 3, 1).reshape(-1, 128)
    assert rel_l2(out, ref) <= TOL


def test_upsample_is_exact(L):
    g = _gen(12)
    x = torch.randn(2, 8, 8, 64, device="cuda", generator=g).half()
    out = L.upsample_nearest2x(x.reshape(-1, 64), batch=2, h=8, w=8)
    ref = F.interpolate(x.permute(0, 3, 1, 2).float(), scale_factor=2.0, mode="nearest").permute(0, 2, 3, 1)
    assert torch.equal(out.reshape(2, 16, 16, 64).float(), ref)


def test_latent_in_out_roundtrip(L):
    """add_noise then pred_original_sample with eps == noise returns the clean latent (scheduler identity)."""
    g = _gen(13)
    x = torch.randn(2, 4, 16, 16, device="cuda", generator=g)
    nz = torch.randn(2, 4, 16, 16, device="cuda", generator=g)
    a, s = 0.9, 0.3
    xin = L.latent_in(x, nz, a, s)
    ref = torch.zeros(2, 256, 64, device="cuda")
    ref[..., :4] = (a * x + s * nz).permute(0, 2, 3, 1).reshape(2, 256, 4)
    assert rel_l2(xin.reshape(2, 256, 64), ref) <= TOL
    eps = torch.zeros(2 * 256, 8, device="cuda", dtype=torch.float16)
    eps[:, :4] = nz.permute(0, 2, 3, 1).reshape(-1, 4).half()
    x0 = L.latent_out(eps, x, nz, a, s)
    assert float((x0 - x).abs().max()) <= 2e-3      # eps was rounded to fp16
    e2 = torch.randn(2 * 256, 8, device="cuda", generator=g).half()
    out = L.latent_out(e2, x, nz, a, s)
    ref = ((a * x + s * nz) - s * e2[:, :4].float().reshape(2, 16, 16, 4).permute(0, 3, 1, 2)) / a
    assert rel_l2(out, ref) <= 1e-5


@pytest.mark.parametrize("cfg,splits", [((1, 5, 1024, False, 4, True, None), 4), ((1, 2, 256, True, 2, True, None), 3),
                                        ((1, 5, 4096, False, 4, True, None), 7), ((1, 2, 512, True, 0, False, None), 2),
                                        ((1, 5, 4096, False, 4, False, None), 0), ((1, 10, 1024, True, 4, True, None), 0)])
def test_shared_attention_split_kv(L, cfg, splits):
    """Split-KV partials + fixed-order combine (explicit split counts and the auto plan for a single identity)."""
    B, H, S, own, n_ref, adain, s_own_o = cfg
    q, kw, ref, _ = _attn_case(L, *cfg, seed=31)
    out = L.shared_attn(q.reshape(-1, H * 64), heads=H, scale=0.125, batch=B, s_q=S, kv_splits=splits, **kw)
    assert rel_l2(out.reshape(B, S, -1), ref) <= TOL
    again = L.shared_attn(q.reshape(-1, H * 64), heads=H, scale=0.125, batch=B, s_q=S, kv_splits=splits, **kw)
    assert torch.equal(out, again)
    unsplit = L.shared_attn(q.reshape(-1, H * 64), heads=H, scale=0.125, batch=B, s_q=S, kv_splits=1, **kw)
    assert rel_l2(out, unsplit) <= 5e-4


def test_shared_attention_large_logits_lazy_rescale(L):
    """Scores that keep growing along the KV axis force the lazy (threshold 2^8) accumulator rescale to fire."""
    g = _gen(32)
    B, H, S, N = 1, 2, 256, 4
    C = H * 64
    q = torch.randn(B, S, C, device="cuda", generator=g).half()
    kr = torch.randn(B, N, S, C, device="cuda", generator=g)
    ramp = torch.linspace(0.2, 6.0, N * S, device="cuda").view(1, N, S, 1)     # later keys get much larger logits
    kr = (kr * ramp).half()
    vr = torch.randn(B, N, S, C, device="cuda", generator=g).half()
    a_s = (torch.rand(B, N, C, device="cuda", generator=g) + 0.5).contiguous()
    a_b = torch.randn(B, N, C, device="cuda", generator=g).contiguous()
    for adain in (False, True):
        vc = [vr[:, r].float() * (a_s[:, r, None] if adain else 1) + (a_b[:, r, None] if adain else 0) for r in range(N)]
        ref, _ = _attn_ref(q, [kr[:, r] for r in range(N)], vc, H, 0.125)
        kw = dict(adain_scale=a_s, adain_shift=a_b) if adain else {}
        out = L.shared_attn(q.reshape(-1, C), heads=H, scale=0.125, batch=B, s_q=S, k_ref=kr.reshape(-1, C),
                            v_ref=vr.reshape(-1, C), n_ref=N, s_ref=S, kv_splits=1, **kw)
        assert rel_l2(out.reshape(B, S, C), ref) <= 2 * TOL
