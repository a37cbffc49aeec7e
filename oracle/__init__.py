"""ORACLE — test infrastructure only.

CPU (plain PyTorch, fp32) restatement of the reference hot path: snap-research/InstantRestore's single-step UNet
forward with the shared-image attention processor and AdaIN. Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this package, and only as the checker or the timed CPU baseline —
never from the product path (instantrestore_b200/), which fails loudly without its CUDA library.

Parity pinning: the reference ships no tests, golden vectors or fixtures, and its arithmetic
lives in un-vendored diffusers==0.24.0 / peft==0.10.0. The oracle is therefore pinned by RUNNING THE REFERENCE'S OWN
unet.py / block.py / attn_processors.py here (oracle/make_golden.py imports them from /root/reference on top of
oracle/shim, a stand-in `diffusers` package that re-exports oracle/diffusers024.py) and committing the resulting
vectors under tests/golden/. The third-party leaf modules themselves (diffusers 0.24.0 Attention, Transformer2DModel,
ResnetBlock2D, ...) remain a restatement of their published semantics: "parity unpinned" for those leaves.
"""
