"""Opt-in by-products of the shared attention for the callers that read them in the reference:

* dense `attention_probs` (B, H, S_q, S_k) — reference attn_processors.py:257-260 keeps the full softmax matrix when
  `save_self_attentions` is set (consumers: test.py:93-108, gradio_demo.py:106-127, coach.py:531-560). The fused kernel
  never materialises it, so this path recomputes it the reference's way: scores = Q_h K_h^T by ir_conv_gemm (fp16, like
  baddbmm under autocast), row softmax in fp32 by ir_softmax_rows. 0.67 GB per 64x64 layer and identity at N=4 —
  diagnostic use only.
* per-chunk attention mass (B, H, n_chunks) comes straight from ir_shared_attn_fwd (`chunk_mass`), no dense matrix.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch

from . import _lib as L

# (tensor [rows, row_width], column offset of head 0, rows of the chunk, row stride between batch entries, first row)
KeyChunk = Tuple[torch.Tensor, int, int, int, int]


def dense_attention_probs(q2d: torch.Tensor, q_col_off: int, chunks: Sequence[KeyChunk], *, batch: int, heads: int,
                          s_q: int, scale: float) -> torch.Tensor:
    s_k = sum(c[2] for c in chunks)
    out = torch.empty((batch, heads, s_q, s_k), dtype=torch.float16, device=q2d.device)
    for b in range(batch):
        for h in range(heads):
            parts: List[torch.Tensor] = []
            for k2d, off, rows, batch_stride, base in chunks:
                r0 = base + b * batch_stride
                parts.append(k2d[r0:r0 + rows, off + h * 64: off + (h + 1) * 64])
            k_h = torch.cat(parts, 0).contiguous()                                   # [S_k, 64]
            q_h = q2d[b * s_q:(b + 1) * s_q, q_col_off + h * 64: q_col_off + (h + 1) * 64]
            L.conv_gemm(q_h, k_h, batch=1, h_in=1, w_in=s_q, c_in=64, out=out[b, h])
            L.softmax_rows(out[b, h], scale)
    return out
