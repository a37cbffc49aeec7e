"""ctypes binding of include/instantrestore_b200.h (the C ABI of the sm_100a kernels).

There is no CPU fallback: if the shared library is missing, loading raises. PyTorch is used only for device memory
and the current CUDA stream; every function below takes tensors, checks dtype/contiguity, and forwards raw pointers.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import torch

_PKG = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("IR_LIB_PATH", _PKG / "libinstantrestore_b200.so"))     # IR_LIB_PATH: A-B a previous build

IR_ACT_NONE, IR_ACT_GEGLU, IR_ACT_SILU = 0, 1, 2

EXPORTED_SYMBOLS = [
    "ir_last_error_string", "ir_version", "ir_check_device", "ir_set_pdl", "ir_launch_count", "ir_conv_gemm", "ir_shared_attn_fwd",
    "ir_shared_attn_workspace_bytes",
    "ir_groupnorm", "ir_groupnorm_workspace_bytes", "ir_groupnorm_fused_supported", "ir_layernorm", "ir_adain_coeffs", "ir_adain_workspace_bytes",
    "ir_concat_freeu", "ir_upsample_nearest2x", "ir_latent_in", "ir_latent_out",
    "ir_softmax_rows", "ir_image_in", "ir_image_in_patches3x3", "ir_image_out", "ir_vae_sample",
    "ir_resample_u8_pass", "ir_u8_to_f16", "ir_image_out_u8",
]


class ConvGemmParams(C.Structure):
    _fields_ = [
        ("a", C.c_void_p), ("batch", C.c_int), ("h_in", C.c_int), ("w_in", C.c_int), ("c_in", C.c_int),
        ("a_row_stride", C.c_int), ("ksize", C.c_int), ("stride", C.c_int),
        ("w", C.c_void_p), ("c_out", C.c_int), ("bias", C.c_void_p), ("residual", C.c_void_p),
        ("res_row_stride", C.c_int), ("act", C.c_int), ("out", C.c_void_p), ("out_row_stride", C.c_int),
        ("tile_n", C.c_int), ("split_k", C.c_int), ("m_sub", C.c_int), ("no_persistent", C.c_int), ("pad_hi_only", C.c_int),
        ("cta_pair", C.c_int), ("gn_partial", C.c_void_p), ("gn_groups", C.c_int), ("halo", C.c_int), ("wide_io", C.c_int),
        ("col_partial", C.c_void_p), ("col_begin", C.c_int), ("upsample2x", C.c_int), ("tma_store", C.c_int),
    ]


class SharedAttnParams(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("q_row_stride", C.c_int), ("q_col_off", C.c_int),
        ("k_own", C.c_void_p), ("v_own", C.c_void_p),
        ("own_row_stride", C.c_int), ("k_own_col_off", C.c_int), ("v_own_col_off", C.c_int),
        ("s_own", C.c_int), ("own_shared", C.c_int),
        ("k_ref", C.c_void_p), ("v_ref", C.c_void_p),
        ("ref_row_stride", C.c_int), ("ref_col_off", C.c_int), ("n_ref", C.c_int), ("s_ref", C.c_int),
        ("adain_scale", C.c_void_p), ("adain_shift", C.c_void_p),
        ("batch", C.c_int), ("heads", C.c_int), ("s_q", C.c_int), ("scale", C.c_float),
        ("out", C.c_void_p), ("out_row_stride", C.c_int), ("chunk_mass", C.c_void_p),
        ("kv_splits", C.c_int), ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
    ]


class GroupNormParams(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("x_row_stride", C.c_int), ("batch", C.c_int), ("hw", C.c_int), ("channels", C.c_int),
        ("groups", C.c_int), ("eps", C.c_float), ("gamma", C.c_void_p), ("beta", C.c_void_p), ("silu", C.c_int),
        ("out", C.c_void_p), ("out_row_stride", C.c_int), ("workspace", C.c_void_p), ("partial_in", C.c_void_p),
        ("fused", C.c_int),
    ]


class LayerNormParams(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("x_row_stride", C.c_int), ("rows", C.c_int), ("channels", C.c_int), ("eps", C.c_float),
        ("gamma", C.c_void_p), ("beta", C.c_void_p), ("out", C.c_void_p), ("out_row_stride", C.c_int),
    ]


class AdainCoeffsParams(C.Structure):
    _fields_ = [
        ("v_own", C.c_void_p), ("own_row_stride", C.c_int), ("v_col_off", C.c_int), ("s_own", C.c_int),
        ("v_ref", C.c_void_p), ("ref_row_stride", C.c_int), ("ref_col_off", C.c_int), ("n_ref", C.c_int),
        ("s_ref", C.c_int), ("batch", C.c_int), ("channels", C.c_int), ("eps", C.c_float),
        ("scale", C.c_void_p), ("shift", C.c_void_p), ("workspace", C.c_void_p),
        ("own_partial", C.c_void_p), ("ref_partial", C.c_void_p),
    ]


class ConcatFreeuParams(C.Structure):
    _fields_ = [
        ("hidden", C.c_void_p), ("skip", C.c_void_p), ("batch", C.c_int), ("h", C.c_int), ("w", C.c_int),
        ("c_hidden", C.c_int), ("c_skip", C.c_int), ("backbone_scale", C.c_float), ("skip_scale", C.c_float),
        ("out", C.c_void_p), ("two_pass", C.c_int),
    ]


_lib = None


def load() -> C.CDLL:
    """Loads the C-ABI library. Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m instantrestore_b200.build` "
            "(or __graft_entry__.build()). There is no CPU/PyTorch fallback for the product path."
        )
    lib = C.CDLL(str(LIB_PATH))
    lib.ir_last_error_string.restype = C.c_char_p
    lib.ir_launch_count.restype = C.c_ulonglong
    lib.ir_groupnorm_workspace_bytes.restype = C.c_size_t
    lib.ir_groupnorm_fused_supported.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int]
    lib.ir_adain_workspace_bytes.restype = C.c_size_t
    lib.ir_shared_attn_workspace_bytes.restype = C.c_size_t
    lib.ir_conv_gemm.argtypes = [C.POINTER(ConvGemmParams), C.c_void_p]
    lib.ir_shared_attn_fwd.argtypes = [C.POINTER(SharedAttnParams), C.c_void_p]
    lib.ir_groupnorm.argtypes = [C.POINTER(GroupNormParams), C.c_void_p]
    lib.ir_layernorm.argtypes = [C.POINTER(LayerNormParams), C.c_void_p]
    lib.ir_adain_coeffs.argtypes = [C.POINTER(AdainCoeffsParams), C.c_void_p]
    lib.ir_concat_freeu.argtypes = [C.POINTER(ConcatFreeuParams), C.c_void_p]
    lib.ir_upsample_nearest2x.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    lib.ir_latent_in.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                 C.c_int, C.c_void_p]
    lib.ir_latent_out.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_int,
                                  C.c_int, C.c_int, C.c_void_p]
    lib.ir_softmax_rows.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p]
    lib.ir_image_in.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    lib.ir_image_in_patches3x3.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    lib.ir_image_out.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                 C.c_int, C.c_void_p]
    lib.ir_vae_sample.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_float, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                  C.c_void_p]
    lib.ir_resample_u8_pass.argtypes = [C.c_void_p, C.c_long, C.c_long, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                        C.c_void_p, C.c_long, C.c_long, C.c_long, C.c_int, C.c_void_p]
    lib.ir_u8_to_f16.argtypes = [C.c_void_p, C.c_long, C.c_long, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    lib.ir_image_out_u8.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    lib.ir_debug_umma.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int] + [C.c_uint] * 6 + [C.c_void_p]
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"{what} failed ({rc}): {load().ir_last_error_string().decode()}")


def set_pdl(enabled: bool) -> bool:
    """Programmatic dependent launch for the launches / graph captures that follow; returns the previous setting."""
    return bool(load().ir_set_pdl(int(bool(enabled))))


def launch_count() -> int:
    return int(load().ir_launch_count())


class Trace:
    """Per-call device timing for the roofline report: while active, every C-ABI call is bracketed by CUDA events
    recorded on the launching stream, together with its algorithmic work (flops, bytes). Not usable under graph
    capture; bench.py runs one eager instrumented step with it.

    The eager event pairs include launch gaps, so the numbers bench.py reports come from `replay_table`: every distinct
    (op, shape) of the step is re-issued `reps` times back to back inside a small CUDA graph (same arguments, same
    buffers, which the trace keeps alive) and timed by replaying that graph — per-launch device time as the kernel
    runs inside the step's own graph, without host launch latency."""
    active: "Trace | None" = None

    def __init__(self):
        self.records = []   # (op, tag, flops, bytes, start_event, end_event, fn, args_without_stream, keep, device)

    def __enter__(self):
        Trace.active = self
        return self

    def __exit__(self, *exc):
        Trace.active = None

    def summary(self):
        torch.cuda.synchronize()
        rows = {}
        for op, tag, flops, nbytes, e0, e1, *_ in self.records:
            r = rows.setdefault((op, tag), dict(op=op, shape=tag, calls=0, ms=0.0, flops=0.0, bytes=0.0))
            r["calls"] += 1
            r["ms"] += e0.elapsed_time(e1)
            r["flops"] += flops
            r["bytes"] += nbytes
        return sorted(rows.values(), key=lambda r: -r["ms"])

    def replay_table(self, reps: int = 8, iters: int = 3):
        """Rows like summary(), with `us` = device time per launch from CUDA-graph replays of `reps` back-to-back
        launches (best of `iters`), `ms` = us * calls."""
        torch.cuda.synchronize()
        rows, first = {}, {}
        for rec in self.records:
            op, tag, flops, nbytes = rec[:4]
            r = rows.setdefault((op, tag), dict(op=op, shape=tag, calls=0, flops=0.0, bytes=0.0, ms_eager=0.0))
            r["calls"] += 1
            r["flops"] += flops
            r["bytes"] += nbytes
            r["ms_eager"] += rec[4].elapsed_time(rec[5])
            first.setdefault((op, tag), rec)
        for key, rec in first.items():
            fn, args, _keep, dev = rec[6:10]
            with on_device(torch.empty(0, device=dev)):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    sp = torch.cuda.current_stream(dev).cuda_stream
                    for _ in range(reps):
                        check(fn(*args, sp), key[0])
                g.replay()
                torch.cuda.synchronize(dev)
                best = float("inf")
                for _ in range(iters):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    g.replay()
                    e1.record()
                    e1.synchronize()
                    best = min(best, e0.elapsed_time(e1))
                del g
            rows[key]["us"] = best * 1e3 / reps
            rows[key]["ms"] = rows[key]["us"] * rows[key]["calls"] * 1e-3
        return sorted(rows.values(), key=lambda r: -r["ms"])


def _run(op: str, tag: str, flops: float, nbytes: float, fn, *args, keep=()) -> None:
    """args[-1] is the stream handle. `keep`: the tensors behind the raw pointers in `args` (kept alive by an active
    Trace so the call can be re-issued by replay_table)."""
    tr = Trace.active
    if tr is None:
        check(fn(*args), op)
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    check(fn(*args), op)
    e1.record()
    dev = next((t.device for t in keep if isinstance(t, torch.Tensor)), torch.device("cuda", torch.cuda.current_device()))
    tr.records.append((op, tag, flops, nbytes, e0, e1, fn, args[:-1], keep, dev))


def stream_ptr(device=None) -> int:
    """Raw handle of torch's current stream on `device` (default: the current device)."""
    return torch.cuda.current_stream(device).cuda_stream


class on_device:
    """Makes the device of tensor `t` current for the duration of a C-ABI call: the library launches on the current
    device (tensor maps, function attributes and the architecture check are per device), so a pipeline built with
    device='cuda:1' must not launch against device 0's context."""
    __slots__ = ("idx", "prev")

    def __init__(self, t: torch.Tensor):
        self.idx = t.device.index if t.device.index is not None else torch.cuda.current_device()

    def __enter__(self):
        self.prev = torch.cuda.current_device()
        if self.prev != self.idx:
            torch.cuda.set_device(self.idx)
        return self

    def __exit__(self, *exc):
        if self.prev != self.idx:
            torch.cuda.set_device(self.prev)


def ptr(t: torch.Tensor | None) -> int | None:
    return None if t is None else t.data_ptr()


def _h(t: torch.Tensor, name: str) -> torch.Tensor:
    if t.dtype != torch.float16 or not t.is_cuda:
        raise TypeError(f"{name}: expected a CUDA fp16 tensor, got {t.dtype} on {t.device}")
    if t.stride(-1) != 1:
        raise ValueError(f"{name}: innermost dimension must be contiguous")
    return t


def _f(t: torch.Tensor | None, name: str) -> torch.Tensor | None:
    if t is None:
        return None
    if t.dtype != torch.float32 or not t.is_cuda or not t.is_contiguous():
        raise TypeError(f"{name}: expected a contiguous CUDA fp32 tensor")
    return t


_scratch_bufs: dict = {}
_scratch_ns = 0


class scratch_namespace:
    """Scratch buffers requested inside this context belong to `ns` (e.g. one CUDA-graph instance): graphs that may
    replay concurrently on different streams must not share split-KV scratch."""

    def __init__(self, ns):
        self.ns = ns

    def __enter__(self):
        global _scratch_ns
        self.prev, _scratch_ns = _scratch_ns, self.ns
        return self

    def __exit__(self, *exc):
        global _scratch_ns
        _scratch_ns = self.prev


def _scratch(device, nbytes: int) -> torch.Tensor:
    """Per-(device, stream) scratch for split-KV partials. Kernels of one stream run in order, so one buffer serves
    every layer launched on it; it only grows, and superseded buffers stay alive because captured CUDA graphs hold their addresses."""
    bufs = _scratch_bufs.setdefault((_scratch_ns, str(device), torch.cuda.current_stream(device).cuda_stream), [])
    if not bufs or bufs[-1].numel() < nbytes:
        bufs.append(torch.empty(nbytes, dtype=torch.uint8, device=device))
    return bufs[-1]


# ------------------------------------------------------------------------------------------------- wrappers
_NARROW_IO = 1 if os.environ.get("IR_WIDE_IO", "1") == "0" else 0      # A-B switch: IR_WIDE_IO=0 -> 128-bit epilogue I/O
_TMA_STORE = 1 if os.environ.get("IR_TMA_STORE", "1") == "0" else 0    # A-B switch: IR_TMA_STORE=0 -> per-thread row stores everywhere


def conv_gemm(a: torch.Tensor, w: torch.Tensor, *, batch: int, h_in: int, w_in: int, c_in: int, ksize: int = 1,
              stride: int = 1, bias: torch.Tensor | None = None, residual: torch.Tensor | None = None,
              act: int = IR_ACT_NONE, out: torch.Tensor | None = None, tile_n: int = 0, split_k: int = 0,
              a_row_stride: int | None = None, pad_hi_only: bool = False, no_persistent: int = 0, m_sub: int = 0, cta_pair: int = 0, halo: int = 0, gn_partial: torch.Tensor | None = None,
              gn_groups: int = 32, wide_io: int = 0, col_partial: torch.Tensor | None = None, col_begin: int = 0,
              upsample2x: bool = False, tma_store: int = 0) -> torch.Tensor:
    """a: fp16 channel-last [batch*h_in*w_in, >=c_in]; w: fp16 [c_out, ksize*ksize*c_in].
    gn_partial: fp32 [batch * (h_out*w_out/32) * gn_groups * 2] (gn_partial_numel) to receive pass A of the next GroupNorm.
    col_partial: fp32 [M/32, c_out - col_begin, 2] to receive per-(32-row slab, column) (mean, M2) of the outputs (AdaIN).
    upsample2x: nearest-2x upsampling + 3x3 convolution as four 2x2 sub-pixel convolutions on the low-resolution input;
    w is then fold_upsample_weight(...) = [4*c_out, 4*c_in] and the result has 4*batch*h_in*w_in rows."""
    _h(a, "a"); _h(w, "w"); _f(bias, "bias"); _f(gn_partial, "gn_partial"); _f(col_partial, "col_partial")
    if upsample2x:
        return _conv_gemm_up2x(a, w, batch=batch, h_in=h_in, w_in=w_in, c_in=c_in, ksize=ksize, stride=stride, bias=bias,
                               out=out, tile_n=tile_n, split_k=split_k, a_row_stride=a_row_stride, no_persistent=no_persistent,
                               cta_pair=cta_pair, gn_partial=gn_partial, gn_groups=gn_groups, wide_io=wide_io)
    c_out = w.shape[0]
    assert w.shape[1] == ksize * ksize * c_in and w.is_contiguous(), (w.shape, ksize, c_in)
    m = batch * (h_in // stride) * (w_in // stride)
    n_out = c_out // 2 if act == IR_ACT_GEGLU else c_out
    if out is None:
        out = torch.empty((m, n_out), dtype=torch.float16, device=a.device)
    p = ConvGemmParams(
        a=ptr(a), batch=batch, h_in=h_in, w_in=w_in, c_in=c_in,
        a_row_stride=a_row_stride if a_row_stride is not None else a.stride(-2),
        ksize=ksize, stride=stride, w=ptr(w), c_out=c_out, bias=ptr(bias),
        residual=ptr(residual), res_row_stride=residual.stride(-2) if residual is not None else 0,
        act=act, out=ptr(out), out_row_stride=out.stride(-2), tile_n=tile_n, split_k=split_k,
        pad_hi_only=int(pad_hi_only), no_persistent=int(no_persistent), m_sub=m_sub, cta_pair=cta_pair, halo=halo, gn_partial=ptr(gn_partial), gn_groups=gn_groups, wide_io=wide_io or _NARROW_IO,
        col_partial=ptr(col_partial), col_begin=col_begin, tma_store=tma_store or _TMA_STORE)
    k_tot = ksize * ksize * c_in
    with on_device(a):
        _run("ir_conv_gemm", f"m{m}_k{k_tot}_n{c_out}_ks{ksize}s{stride}", 2.0 * m * k_tot * c_out,
             2.0 * (m * c_in * (1 if ksize == 1 else stride * stride) + c_out * k_tot + m * n_out
                    + (m * n_out if residual is not None else 0)),
             load().ir_conv_gemm, C.byref(p), stream_ptr(a.device), keep=(a, w, bias, residual, out, gn_partial, col_partial))
    return out


def _conv_gemm_up2x(a, w, *, batch, h_in, w_in, c_in, ksize, stride, bias, out, tile_n, split_k, a_row_stride, no_persistent,
                    cta_pair, gn_partial, gn_groups, wide_io):
    assert ksize == 3 and stride == 1, "upsample2x folds into a 3x3 stride-1 convolution"
    assert w.shape[0] % 4 == 0 and w.shape[1] == 4 * c_in and w.is_contiguous(), (w.shape, c_in)
    c_out = w.shape[0] // 4
    m = 4 * batch * h_in * w_in                        # output pixels
    if out is None:
        out = torch.empty((m, c_out), dtype=torch.float16, device=a.device)
    p = ConvGemmParams(
        a=ptr(a), batch=batch, h_in=h_in, w_in=w_in, c_in=c_in,
        a_row_stride=a_row_stride if a_row_stride is not None else a.stride(-2),
        ksize=3, stride=1, w=ptr(w), c_out=c_out, bias=ptr(bias), residual=None, res_row_stride=0,
        act=IR_ACT_NONE, out=ptr(out), out_row_stride=out.stride(-2), tile_n=tile_n, split_k=split_k,
        pad_hi_only=0, no_persistent=int(no_persistent), m_sub=0, cta_pair=cta_pair, halo=0, gn_partial=ptr(gn_partial),
        gn_groups=gn_groups, wide_io=wide_io or _NARROW_IO, col_partial=None, col_begin=0, upsample2x=1)
    k_tot = 4 * c_in                                   # executed multiply-adds: 4/9 of the 3x3 convolution on the upsampled tensor
    with on_device(a):
        _run("ir_conv_gemm", f"m{m}_k{k_tot}_n{c_out}_ks3s1_up2x", 2.0 * m * k_tot * c_out,
             2.0 * (m // 4 * c_in + 4 * c_out * k_tot + m * c_out),
             load().ir_conv_gemm, C.byref(p), stream_ptr(a.device), keep=(a, w, bias, out, gn_partial))
    return out


def shared_attn(q: torch.Tensor, *, heads: int, scale: float, batch: int, s_q: int,
                k_own: torch.Tensor | None = None, v_own: torch.Tensor | None = None, s_own: int = 0,
                own_shared: bool = False, q_col_off: int = 0, k_own_col_off: int = 0, v_own_col_off: int = 0,
                k_ref: torch.Tensor | None = None, v_ref: torch.Tensor | None = None, n_ref: int = 0, s_ref: int = 0,
                ref_col_off: int = 0, adain_scale: torch.Tensor | None = None, adain_shift: torch.Tensor | None = None,
                out: torch.Tensor | None = None, kv_splits: int = 0, chunk_mass: bool = False):
    """q: fp16 [batch*s_q, row]; k_own/v_own: fp16 [(batch|1)*s_own, row]; k_ref/v_ref: fp16 [batch*n_ref*s_ref, row]."""
    _h(q, "q")
    if out is None:
        out = torch.empty((batch * s_q, heads * 64), dtype=torch.float16, device=q.device)
    p = SharedAttnParams(
        q=ptr(q), q_row_stride=q.stride(-2), q_col_off=q_col_off,
        k_own=ptr(k_own), v_own=ptr(v_own), own_row_stride=k_own.stride(-2) if k_own is not None else 0,
        k_own_col_off=k_own_col_off, v_own_col_off=v_own_col_off, s_own=s_own, own_shared=int(own_shared),
        k_ref=ptr(k_ref), v_ref=ptr(v_ref), ref_row_stride=k_ref.stride(-2) if k_ref is not None else 0,
        ref_col_off=ref_col_off, n_ref=n_ref, s_ref=s_ref,
        adain_scale=ptr(_f(adain_scale, "adain_scale")), adain_shift=ptr(_f(adain_shift, "adain_shift")),
        batch=batch, heads=heads, s_q=s_q, scale=scale, out=ptr(out), out_row_stride=out.stride(-2), chunk_mass=None,
        kv_splits=kv_splits)
    mass = None
    if chunk_mass:
        n_chunks = (1 if k_own is not None else 0) + n_ref
        mass = torch.empty((batch, heads, n_chunks), dtype=torch.float32, device=q.device)
        ws = _scratch(q.device, batch * heads * ((s_q + 255) // 256) * 256 * (n_chunks + 1) * 8)
        p.chunk_mass, p.workspace, p.workspace_bytes, p.kv_splits = ptr(mass), ptr(ws), ws.numel(), 1
    else:
        ws_bytes = load().ir_shared_attn_workspace_bytes(batch, heads, s_q) if kv_splits != 1 else 0
        if ws_bytes:
            ws = _scratch(q.device, ws_bytes)
            p.workspace, p.workspace_bytes = ptr(ws), ws.numel()
    if k_own is not None:
        _h(k_own, "k_own"); _h(v_own, "v_own")
        assert k_own.stride(-2) == v_own.stride(-2)
    if k_ref is not None:
        _h(k_ref, "k_ref"); _h(v_ref, "v_ref")
        assert k_ref.stride(-2) == v_ref.stride(-2)
    s_kv = (s_own if k_own is not None else 0) + n_ref * s_ref
    with on_device(q):
        _run("ir_shared_attn_fwd", f"b{batch}_h{heads}_sq{s_q}_skv{s_kv}" + ("_adain" if adain_scale is not None else ""), 4.0 * batch * heads * s_q * s_kv * 64,
             2.0 * batch * heads * 64 * (2 * s_q + 2 * s_kv), load().ir_shared_attn_fwd, C.byref(p), stream_ptr(q.device), keep=(q, k_own, v_own, k_ref, v_ref, adain_scale, adain_shift, out, mass, p.workspace and ws))
    return (out, mass) if chunk_mass else out


def gn_partial_supported(hw: int, channels: int, groups: int = 32) -> bool:
    """Shapes for which conv_gemm can emit GroupNorm pass A of its output (whole groups per 32-column accumulator chunk,
    whole 32-pixel slabs per image)."""
    return channels % groups == 0 and channels // groups in (4, 8, 16) and hw % 128 == 0


def gn_partial_numel(batch: int, hw: int, groups: int = 32) -> int:
    return batch * (hw // 32) * groups * 2


_GN_FUSED_MODE = {"0": 1, "2": 2}.get(os.environ.get("IR_GN_SINGLE_LAUNCH", "1"), 0)   # A-B switch: 0 -> three-kernel path


def gn_fused_supported(batch: int, hw: int, channels: int, groups: int = 32) -> bool:
    """True when ir_groupnorm runs as ONE launch with ONE read of x (cluster kernel); conv_gemm's gn_partial is then not
    needed for this tensor."""
    return _GN_FUSED_MODE != 1 and bool(load().ir_groupnorm_fused_supported(batch, hw, channels, groups))


def groupnorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, *, batch: int, hw: int, groups: int = 32,
              eps: float = 1e-5, silu: bool = False, out: torch.Tensor | None = None,
              workspace: torch.Tensor | None = None, partial_in: torch.Tensor | None = None, fused: int | None = None) -> torch.Tensor:
    """partial_in: the gn_partial tensor a conv_gemm call filled for exactly this x (pass A is then skipped).
    fused: None = library default (single-launch cluster kernel when supported), 1 = three-kernel path, 2 = require it."""
    _h(x, "x"); _f(gamma, "gamma"); _f(beta, "beta"); _f(partial_in, "partial_in")
    channels = x.shape[-1]
    fused = _GN_FUSED_MODE if fused is None else fused
    if out is None:
        out = torch.empty((batch * hw, channels), dtype=torch.float16, device=x.device)
    single = fused != 1 and bool(load().ir_groupnorm_fused_supported(batch, hw, channels, groups))
    if workspace is None and not single:
        workspace = torch.empty(load().ir_groupnorm_workspace_bytes(batch, groups) // 4, dtype=torch.float32,
                                device=x.device)
    p = GroupNormParams(x=ptr(x), x_row_stride=x.stride(-2), batch=batch, hw=hw, channels=channels, groups=groups,
                        eps=eps, gamma=ptr(gamma), beta=ptr(beta), silu=int(silu), out=ptr(out),
                        out_row_stride=out.stride(-2), workspace=ptr(workspace), partial_in=ptr(partial_in), fused=fused)
    with on_device(x):
        _run("ir_groupnorm", f"b{batch}_hw{hw}_c{channels}" + ("_1launch" if single else ""), 0.0, 4.0 * batch * hw * channels, load().ir_groupnorm,
             C.byref(p), stream_ptr(x.device), keep=(x, gamma, beta, out, workspace, partial_in))
    return out


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, *, eps: float = 1e-5,
              out: torch.Tensor | None = None) -> torch.Tensor:
    _h(x, "x"); _f(gamma, "gamma"); _f(beta, "beta")
    rows, channels = x.shape[-2], x.shape[-1]
    if out is None:
        out = torch.empty((rows, channels), dtype=torch.float16, device=x.device)
    p = LayerNormParams(x=ptr(x), x_row_stride=x.stride(-2), rows=rows, channels=channels, eps=eps, gamma=ptr(gamma),
                        beta=ptr(beta), out=ptr(out), out_row_stride=out.stride(-2))
    with on_device(x):
        _run("ir_layernorm", f"r{rows}_c{channels}", 0.0, 4.0 * rows * channels, load().ir_layernorm, C.byref(p),
             stream_ptr(x.device), keep=(x, gamma, beta, out))
    return out


def col_partial_supported(m: int, c_out: int, col_begin: int) -> bool:
    """Shapes for which conv_gemm can emit the per-slab column moments of its outputs (AdaIN statistics)."""
    return m % 128 == 0 and c_out % 32 == 0 and col_begin % 32 == 0


def adain_coeffs(v_own: torch.Tensor | None, v_ref: torch.Tensor | None, *, batch: int, s_own: int, n_ref: int, s_ref: int,
                 channels: int, v_col_off: int = 0, ref_col_off: int = 0, eps: float = 1e-5,
                 scale: torch.Tensor | None = None, shift: torch.Tensor | None = None,
                 workspace: torch.Tensor | None = None, own_partial: torch.Tensor | None = None,
                 ref_partial: torch.Tensor | None = None):
    """own_partial / ref_partial: the col_partial tensors the QKV GEMMs filled for the own / reference V ([batch*s/32, C, 2]
    and [batch*n_ref*s/32, C, 2]); v_own / v_ref are then not read (may be None)."""
    if own_partial is not None:
        _f(own_partial, "own_partial"); _f(ref_partial, "ref_partial")
        dev = own_partial.device
        if scale is None:
            scale = torch.empty((batch, n_ref, channels), dtype=torch.float32, device=dev)
        if shift is None:
            shift = torch.empty((batch, n_ref, channels), dtype=torch.float32, device=dev)
        p = AdainCoeffsParams(v_own=None, own_row_stride=0, v_col_off=0, s_own=s_own, v_ref=None, ref_row_stride=0, ref_col_off=0,
                              n_ref=n_ref, s_ref=s_ref, batch=batch, channels=channels, eps=eps, scale=ptr(scale), shift=ptr(shift),
                              workspace=None, own_partial=ptr(own_partial), ref_partial=ptr(ref_partial))
        with on_device(own_partial):
            _run("ir_adain_coeffs", f"b{batch}_n{n_ref}_s{s_ref}_c{channels}_partials", 0.0,
                 8.0 * batch * channels * (s_own + n_ref * s_ref) / 32, load().ir_adain_coeffs, C.byref(p), stream_ptr(dev),
                 keep=(own_partial, ref_partial, scale, shift))
        return scale, shift
    _h(v_own, "v_own"); _h(v_ref, "v_ref")
    dev = v_own.device
    if scale is None:
        scale = torch.empty((batch, n_ref, channels), dtype=torch.float32, device=dev)
    if shift is None:
        shift = torch.empty((batch, n_ref, channels), dtype=torch.float32, device=dev)
    if workspace is None:
        workspace = torch.empty(load().ir_adain_workspace_bytes(batch, n_ref, channels) // 4, dtype=torch.float32,
                                device=dev)
    p = AdainCoeffsParams(v_own=ptr(v_own), own_row_stride=v_own.stride(-2), v_col_off=v_col_off, s_own=s_own,
                          v_ref=ptr(v_ref), ref_row_stride=v_ref.stride(-2), ref_col_off=ref_col_off, n_ref=n_ref,
                          s_ref=s_ref, batch=batch, channels=channels, eps=eps, scale=ptr(scale), shift=ptr(shift),
                          workspace=ptr(workspace))
    with on_device(v_own):
        _run("ir_adain_coeffs", f"b{batch}_n{n_ref}_s{s_ref}_c{channels}", 0.0,
             2.0 * batch * channels * (s_own + n_ref * s_ref), load().ir_adain_coeffs, C.byref(p), stream_ptr(v_own.device), keep=(v_own, v_ref, scale, shift, workspace))
    return scale, shift


def concat_freeu(hidden: torch.Tensor, skip: torch.Tensor, *, batch: int, h: int, w: int, backbone_scale: float = 1.0,
                 skip_scale: float = 1.0, out: torch.Tensor | None = None, two_pass: bool = False) -> torch.Tensor:
    _h(hidden, "hidden"); _h(skip, "skip")
    assert hidden.is_contiguous() and skip.is_contiguous()
    ch, cs = hidden.shape[-1], skip.shape[-1]
    if out is None:
        out = torch.empty((batch * h * w, ch + cs), dtype=torch.float16, device=hidden.device)
    p = ConcatFreeuParams(hidden=ptr(hidden), skip=ptr(skip), batch=batch, h=h, w=w, c_hidden=ch, c_skip=cs,
                          backbone_scale=backbone_scale, skip_scale=skip_scale, out=ptr(out), two_pass=int(two_pass))
    with on_device(hidden):
        _run("ir_concat_freeu", f"b{batch}_hw{h * w}_c{ch}+{cs}", 0.0, 4.0 * batch * h * w * (ch + cs),
             load().ir_concat_freeu, C.byref(p), stream_ptr(hidden.device), keep=(hidden, skip, out))
    return out


def upsample_nearest2x(x: torch.Tensor, *, batch: int, h: int, w: int, out: torch.Tensor | None = None) -> torch.Tensor:
    _h(x, "x")
    assert x.is_contiguous()
    c = x.shape[-1]
    if out is None:
        out = torch.empty((batch * 4 * h * w, c), dtype=torch.float16, device=x.device)
    with on_device(x):
        _run("ir_upsample_nearest2x", f"b{batch}_hw{h * w}_c{c}", 0.0, 10.0 * batch * h * w * c,
             load().ir_upsample_nearest2x, ptr(x), ptr(out), batch, h, w, c, stream_ptr(x.device), keep=(x, out))
    return out


def latent_in(x: torch.Tensor, noise: torch.Tensor | None, a: float, s: float, *, c_pad: int = 64,
              out: torch.Tensor | None = None) -> torch.Tensor:
    """x, noise: fp32 NCHW -> fp16 channel-last [B*HW, c_pad] = a*x + s*noise (zero-padded channels)."""
    _f(x, "x"); _f(noise, "noise")
    b, c, hh, ww = x.shape
    if out is None:
        out = torch.empty((b * hh * ww, c_pad), dtype=torch.float16, device=x.device)
    with on_device(x):
        check(load().ir_latent_in(ptr(x), ptr(noise), a, s, ptr(out), b, c, hh * ww, c_pad, stream_ptr(x.device)), "ir_latent_in")
    return out


def latent_out(eps: torch.Tensor, x: torch.Tensor, noise: torch.Tensor | None, a: float, s: float, *,
               out: torch.Tensor | None = None) -> torch.Tensor:
    """eps: fp16 channel-last [B*HW, >=C]; x, noise: fp32 NCHW -> fp32 NCHW ((a*x + s*noise) - s*eps) / a."""
    _h(eps, "eps"); _f(x, "x"); _f(noise, "noise")
    b, c, hh, ww = x.shape
    if out is None:
        out = torch.empty_like(x)
    with on_device(eps):
        check(load().ir_latent_out(ptr(eps), eps.stride(-2), ptr(x), ptr(noise), a, s, ptr(out), b, c, hh * ww,
                                   stream_ptr(eps.device)), "ir_latent_out")
    return out


def softmax_rows(x: torch.Tensor, scale: float) -> torch.Tensor:
    """In-place softmax(x * scale) over the last dim of an fp16 [rows, cols] matrix (fp32 math)."""
    _h(x, "x")
    rows, cols = x.shape
    with on_device(x):
        _run("ir_softmax_rows", f"r{rows}_c{cols}", 0.0, 4.0 * rows * cols, load().ir_softmax_rows, ptr(x), rows, cols,
             x.stride(0), scale, stream_ptr(x.device), keep=(x,))
    return x


def image_in(x: torch.Tensor, *, c_pad: int = 64, out: torch.Tensor | None = None) -> torch.Tensor:
    """x: fp16/fp32 NCHW image batch -> fp16 channel-last [B*H*W, c_pad] (zero-padded channels)."""
    if not x.is_cuda or not x.is_contiguous() or x.dtype not in (torch.float16, torch.float32):
        raise TypeError("image_in: expected a contiguous CUDA fp16/fp32 NCHW tensor")
    b, c, hh, ww = x.shape
    if out is None:
        out = torch.empty((b * hh * ww, c_pad), dtype=torch.float16, device=x.device)
    with on_device(x):
        _run("ir_image_in", f"b{b}_hw{hh * ww}", 0.0, b * hh * ww * (c * x.element_size() + 2.0 * c_pad), load().ir_image_in,
             ptr(x), int(x.dtype == torch.float32), ptr(out), b, c, hh * ww, c_pad, stream_ptr(x.device), keep=(x, out))
    return out


def image_in_patches3x3(x: torch.Tensor, *, out: torch.Tensor | None = None) -> torch.Tensor:
    """x: fp16/fp32 NCHW image batch (9 * channels <= 64) -> fp16 [B*H*W, 64]: the 3x3 patch of every pixel, (ky, kx, ch)
    order, zero padded — the A operand of conv_in as a K = 64 GEMM (weights: weights.patch_conv_weight)."""
    if not x.is_cuda or not x.is_contiguous() or x.dtype not in (torch.float16, torch.float32):
        raise TypeError("image_in_patches3x3: expected a contiguous CUDA fp16/fp32 NCHW tensor")
    b, c, hh, ww = x.shape
    if out is None:
        out = torch.empty((b * hh * ww, 64), dtype=torch.float16, device=x.device)
    with on_device(x):
        _run("ir_image_in_patches3x3", f"b{b}_hw{hh * ww}", 0.0, b * hh * ww * (c * x.element_size() + 2.0 * 64), load().ir_image_in_patches3x3,
             ptr(x), int(x.dtype == torch.float32), ptr(out), b, c, hh, ww, stream_ptr(x.device), keep=(x, out))
    return out


def image_out(y: torch.Tensor, *, batch: int, c: int, h: int, w: int, lo: float = -1.0, hi: float = 1.0,
              dtype: torch.dtype = torch.float16, out: torch.Tensor | None = None) -> torch.Tensor:
    """y: fp16 channel-last [B*H*W, >=c] -> NCHW clamp(lo, hi)."""
    _h(y, "y")
    if out is None:
        out = torch.empty((batch, c, h, w), dtype=dtype, device=y.device)
    with on_device(y):
        _run("ir_image_out", f"b{batch}_hw{h * w}", 0.0, batch * h * w * c * (2.0 + out.element_size()), load().ir_image_out,
             ptr(y), y.stride(-2), lo, hi, ptr(out), int(out.dtype == torch.float32), batch, c, h * w, stream_ptr(y.device), keep=(y, out))
    return out


def vae_sample(moments: torch.Tensor, eps: torch.Tensor | None, scale: float, *, batch: int, c: int, h: int, w: int,
               out: torch.Tensor | None = None) -> torch.Tensor:
    """moments: fp16 channel-last [B*H*W, >=2c] (mean | logvar); eps fp32 NCHW or None -> fp32 NCHW latent * scale."""
    _h(moments, "moments"); _f(eps, "eps")
    if out is None:
        out = torch.empty((batch, c, h, w), dtype=torch.float32, device=moments.device)
    with on_device(moments):
        check(load().ir_vae_sample(ptr(moments), moments.stride(-2), ptr(eps), scale, ptr(out), batch, c, h * w, stream_ptr(moments.device)),
              "ir_vae_sample")
    return out


def resample_u8_pass(src: torch.Tensor, src_off: int, in_stride_axis: int, in_stride_other: int, n_out: int, n_other: int,
                     bounds: torch.Tensor, kk: torch.Tensor, first: int, out: torch.Tensor, out_stride_axis: int,
                     out_stride_other: int, out_stride_c: int, f16_norm: bool) -> None:
    """One pass of Pillow's 8-bit resampling (see include/instantrestore_b200.h). `src`: CUDA uint8, read from byte
    offset `src_off`; `bounds` int32 [n, 2], `kk` int32 [n, ksize] on the device."""
    assert src.dtype == torch.uint8 and src.is_cuda and bounds.dtype == torch.int32 and kk.dtype == torch.int32
    with on_device(src):
        _run("ir_resample_u8_pass", f"o{n_out}_j{n_other}_k{kk.shape[1]}", 0.0, 3.0 * n_out * n_other * (kk.shape[1] + (2 if f16_norm else 1)),
             load().ir_resample_u8_pass, src.data_ptr() + src_off, in_stride_axis, in_stride_other, n_out, n_other, ptr(bounds), ptr(kk),
             kk.shape[1], first, ptr(out), out_stride_axis, out_stride_other, out_stride_c, int(f16_norm), stream_ptr(src.device),
             keep=(src, bounds, kk, out))


def u8_to_f16(src: torch.Tensor, src_off: int, stride_y: int, stride_x: int, h: int, w: int, out: torch.Tensor) -> None:
    """uint8 window [h, w, 3] (byte strides, starting at byte src_off) -> normalised fp16 NCHW [3, h, w]."""
    assert src.dtype == torch.uint8 and src.is_cuda and out.dtype == torch.float16
    with on_device(src):
        _run("ir_u8_to_f16", f"hw{h * w}", 0.0, 9.0 * h * w, load().ir_u8_to_f16, src.data_ptr() + src_off, stride_y, stride_x, h, w,
             ptr(out), stream_ptr(src.device), keep=(src, out))


def image_out_u8(pred: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """tensor2im(unnorm=True) of the reference on the GPU: fp16 NCHW prediction [B, 3, H, W] -> uint8 [B, H, W, 3]."""
    if pred.dtype != torch.float16 or not pred.is_cuda or not pred.is_contiguous() or pred.shape[1] != 3:
        raise TypeError("image_out_u8: expected a contiguous CUDA fp16 [B, 3, H, W] tensor")
    b, _, hh, ww = pred.shape
    if out is None:
        out = torch.empty((b, hh, ww, 3), dtype=torch.uint8, device=pred.device)
    with on_device(pred):
        _run("ir_image_out_u8", f"b{b}_hw{hh * ww}", 0.0, 9.0 * b * hh * ww, load().ir_image_out_u8, ptr(pred), ptr(out), b, hh * ww,
             stream_ptr(pred.device), keep=(pred, out))
    return out
