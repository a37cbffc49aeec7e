"""Generates tests/golden/*.npz by running the REFERENCE'S OWN code (face_replace/models/unet_2d_condition/unet.py,
block.py and face_replace/models/attn_processors.py, imported from /root/reference) on top of oracle/shim.
Runs only in the build container (the GPU box has no /root/reference); the vectors it writes are committed.

    python -m oracle.make_golden            # all cases
    python -m oracle.make_golden --no-full  # skip the full-width SD-Turbo case (about a minute of CPU)

Each file stores the seeds/config needed to rebuild the inputs from oracle/synth.py plus the reference outputs
(fp32). Inputs are NOT stored (they are regenerated from the seeds), outputs are small (<= 64 KB each).
"""
from __future__ import annotations

import argparse
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
GOLDEN = ROOT / "tests" / "golden"
REFERENCE = Path("/root/reference")

ATTN_CASES = [  # name, heads, S, n_ref, use_adain, train_input, zeroed_slots
    ("attn_plain", 2, 64, 0, False, True, 0),
    ("attn_refs_only", 2, 64, 2, False, False, 0),
    ("attn_own_plus_refs", 2, 64, 2, False, True, 0),
    ("attn_adain_refs_only", 2, 64, 3, True, False, 0),
    ("attn_adain_own_plus_refs", 1, 128, 2, True, True, 0),
    ("attn_adain_padded_slot", 2, 64, 3, True, False, 1),
]
UNET_CASES = [  # name, batch, n_ref, use_adain, train_input, lora_rank, valid
    ("unet_tiny_base", 1, 2, False, True, 0, None),
    ("unet_tiny_final", 2, 2, True, False, 4, None),
    ("unet_tiny_padded", 2, 3, True, False, 4, [3, 1]),
    # reference-count sweep of BASELINE configs[4] (N_ref in {1, 2, 4, 8}); N = 2, 3 above, N = 4 at full width
    ("unet_tiny_n1", 2, 1, True, False, 4, None),
    ("unet_tiny_n8", 1, 8, True, False, 0, None),
    ("unet_tiny_n8_ragged_own", 2, 8, True, True, 4, [8, 5]),
]


# Full-width cases of the BENCHMARKED configurations (BASELINE.json configs[1..4]); generated with --full-only / default.
# The shared-attention layer at SD-Turbo widths: (heads, S) = (20, 256), (10, 1024), (5, 4096), N_ref in {1, 2, 8}
# (N_ref = 4 is inside the full pipeline cases). Output rows are subsampled (every `row_step`-th token) to keep the
# fixtures small; inputs are rebuilt from the seed.
ATTN_FULL_CASES = [  # name, heads, S, n_ref, use_adain, train_input, zeroed_slots, row_step
    (f"attn_full_s{s}_n{n}", h, s, n, True, False, (2 if n == 8 else 0), max(1, s // 64))
    for (h, s) in [(20, 256), (10, 1024), (5, 4096)] for n in (1, 2, 8)
] + [("attn_full_s4096_n4_own", 5, 4096, 4, True, True, 0, 64)]
FULL_LATENT_CASES = [  # name, batch, n_ref, use_adain, train_input, lora_rank, valid
    ("unet_full_final_n1", 1, 1, True, False, 4, None),
    ("unet_full_final_n2", 1, 2, True, False, 4, None),
    ("unet_full_final_n8", 1, 8, True, False, 4, None),
    ("unet_full_final_b8_n4", 8, 4, True, False, 4, [4, 4, 3, 4, 1, 4, 4, 2]),
]
# whole image pipeline at the benchmarked geometry: 512 x 512 images, SD-Turbo + sd-vae-ft-mse widths
IMAGE_FULL_CASES = [  # name, batch, n_ref, use_adain, train_input, lora_rank_unet, lora_rank_vae, use_shortcuts
    ("image_full_final_n4", 1, 4, True, False, 4, 4, False),
]


def full_image_models(use_adain, train_input, lora_unet, lora_vae, use_shortcuts, reference_forwards, RefUNet=None, ref_ap=None):
    """Full-geometry models (SD-Turbo UNet widths, sd-vae-ft-mse VAE widths) for the 512 x 512 image cases."""
    from oracle import synth
    from oracle.pipeline import ImageRestorePipeline, LatentRestorePipeline
    from oracle.unet import UNetConfig
    from oracle.vae import VaeConfig
    ucfg = UNetConfig()
    vcfg = VaeConfig(use_shortcuts=use_shortcuts)
    flags = synth.ModelFlags(use_adain=use_adain, train_input=train_input)
    if RefUNet is not None:
        latent = build_pipeline(RefUNet, ref_ap, ucfg, flags, lora_unet)
    else:
        latent = LatentRestorePipeline(synth.make_unet(ucfg, seed=0, lora_rank=lora_unet), synth.make_unet(ucfg, seed=0),
                                       synth.caption_embedding(ucfg.cross_attention_dim), flags)
    vae = synth.make_vae(vcfg, seed=100, lora_rank=lora_vae)
    ovae = synth.make_vae(VaeConfig(), seed=100)
    if reference_forwards:
        bind_reference_vae_forwards(vae)
        bind_reference_vae_forwards(ovae)
        ovae.decoder.ignore_skip = True
    return ImageRestorePipeline(latent, vae, ovae)


def import_reference():
    if not REFERENCE.exists():
        raise RuntimeError("/root/reference is not mounted: golden vectors can only be regenerated in the build container")
    # this repo ships a `face_replace` package of its own (the drop-in import paths): make sure the REFERENCE's is the one
    # imported here — its directory goes first on sys.path and any already-imported copy is dropped
    for mod in [m for m in sys.modules if m == "face_replace" or m.startswith("face_replace.")]:
        del sys.modules[mod]
    sys.path.insert(0, str(ROOT / "oracle" / "shim"))
    sys.path.insert(0, str(REFERENCE))
    from face_replace.models import attn_processors as ref_ap
    if not str(Path(ref_ap.__file__).resolve()).startswith(str(REFERENCE)):
        raise RuntimeError(f"expected the reference's attn_processors, imported {ref_ap.__file__}")
    from face_replace.models.unet_2d_condition import block as ref_block
    from face_replace.models.unet_2d_condition.unet import UNet2DConditionModel as RefUNet
    return ref_ap, ref_block, RefUNet


def attn_inputs(heads, s, n_ref, zeroed, seed=7, batch=2):
    """Inputs of one processor call: hidden states, an Attention module, reference keys/values (B, N, S, C)."""
    from oracle.diffusers024 import Attention
    from oracle.synth import seeded_init_
    g = torch.Generator().manual_seed(seed)
    c = heads * 64
    attn = seeded_init_(Attention(query_dim=c, heads=heads, dim_head=64), seed).eval().requires_grad_(False)
    hidden = torch.randn(batch, s, c, generator=g)
    rk = torch.randn(batch, max(n_ref, 1), s, c, generator=g)
    rv = torch.randn(batch, max(n_ref, 1), s, c, generator=g) * 1.5 + 0.25
    if zeroed:
        rk[-1, -zeroed:] = 0
        rv[-1, -zeroed:] = 0
    return attn, hidden, rk, rv


def faceid_case(proc_cls, device="cpu"):
    """FaceIDAttnProcessor(hidden 128, cross_attention_dim 256, embed_dim 512) on 4 face embeddings per sample."""
    from oracle.diffusers024 import Attention
    from oracle.synth import seeded_init_
    g = torch.Generator().manual_seed(17)
    attn = seeded_init_(Attention(query_dim=128, cross_attention_dim=256, heads=2, dim_head=64), 17).eval().requires_grad_(False)
    proc = seeded_init_(proc_cls(hidden_size=128, cross_attention_dim=256, embed_dim=512), 18).eval().requires_grad_(False)
    hidden = torch.randn(2, 64, 128, generator=g)
    faces = torch.randn(2, 4, 512, generator=g)
    attn, proc = attn.to(device), proc.to(device)
    with torch.no_grad():
        out = proc(attn, hidden.to(device), encoder_hidden_states=faces.to(device))
    return out, proc


def build_pipeline(RefUNet, ref_ap, cfg, flags, lora_rank, seed=0):
    """Reference UNet classes + reference processors, weights copied from the seeded oracle models."""
    from oracle import synth
    from oracle.pipeline import LatentRestorePipeline
    kw = dict(sample_size=cfg.sample_size, block_out_channels=cfg.block_out_channels,
              attention_head_dim=cfg.attention_head_dim, cross_attention_dim=cfg.cross_attention_dim,
              use_linear_projection=True)
    ref_unet, ref_orig = RefUNet(**kw), RefUNet(**kw)
    o_unet = synth.make_unet(cfg, seed=seed, lora_rank=lora_rank)
    o_orig = synth.make_unet(cfg, seed=seed)
    if lora_rank:
        from oracle.diffusers024 import add_lora
        add_lora(ref_unet, synth.UNET_LORA_TARGETS, r=lora_rank, alpha=lora_rank // 2)
    ref_unet.load_state_dict(o_unet.state_dict(), strict=True)
    ref_orig.load_state_dict(o_orig.state_dict(), strict=True)
    for m in (ref_unet, ref_orig):
        m.enable_freeu(0.9, 0.2, 1.4, 1.6)
        m.eval().requires_grad_(False)
    cap = synth.caption_embedding(cfg.cross_attention_dim)
    return LatentRestorePipeline(ref_unet, ref_orig, cap, flags, processors=ref_ap)


FACEID_CASE = ("unet_tiny_faceid", 2, 3, True, False, 4)     # name, batch, n_ref, use_adain, train_input, lora_rank


def faceid_pipeline_case(RefUNet, ref_ap, write=True):
    from oracle import synth
    from oracle.unet import UNetConfig
    name, batch, n_ref, use_adain, train_input, lora_rank = FACEID_CASE
    tiny = UNetConfig.tiny()
    flags = synth.ModelFlags(use_adain=use_adain, train_input=train_input, condition_on_face_embeds=True)
    pipe = build_pipeline(RefUNet, ref_ap, tiny, flags, lora_rank)
    synth.seed_face_processors(pipe.unet)
    enc, refs, nm, nr = synth.latents(batch, n_ref, tiny.sample_size)
    out = pipe.forward_latents(enc, refs, nm, nr, face_embeds=synth.face_embeddings(batch))
    if write:
        np.savez_compressed(GOLDEN / f"{name}.npz", x0=out.numpy(), meta=np.array([batch, n_ref, int(use_adain), int(train_input), lora_rank]))
        print(name, tuple(out.shape), float(out.std()))
    return out


VAE_CASES = [  # name, use_shortcuts, lora_rank
    ("vae_tiny_plain", False, 0),
    ("vae_tiny_lora", False, 4),
    ("vae_tiny_shortcuts", True, 4),
]
IMAGE_CASES = [  # name, batch, n_ref, use_adain, train_input, lora_rank_unet, lora_rank_vae, use_shortcuts
    ("image_tiny_final", 1, 2, True, False, 4, 4, False),
    ("image_tiny_shortcuts", 2, 2, False, True, 0, 4, True),
]
IMAGE_SIZE, IMAGE_LATENT = 128, 16


def bind_reference_vae_forwards(vae):
    """Replaces the oracle's encoder/decoder forwards with the REFERENCE'S OWN face_replace/models/model.py
    my_vae_encoder_fwd / my_vae_decoder_fwd (what pix2pix_turbo.py:40-41 does)."""
    from face_replace.models import model as ref_model
    vae.encoder.forward = ref_model.my_vae_encoder_fwd.__get__(vae.encoder, vae.encoder.__class__)
    vae.decoder.forward = ref_model.my_vae_decoder_fwd.__get__(vae.decoder, vae.decoder.__class__)
    return vae


def tiny_image_models(use_adain, train_input, lora_unet, lora_vae, use_shortcuts, reference_forwards, RefUNet=None, ref_ap=None):
    from oracle import synth
    from oracle.pipeline import ImageRestorePipeline, LatentRestorePipeline
    from oracle.unet import UNetConfig
    from oracle.vae import VaeConfig
    ucfg = UNetConfig.tiny(sample_size=IMAGE_LATENT)
    vcfg = VaeConfig.tiny()
    vcfg.use_shortcuts = use_shortcuts
    flags = synth.ModelFlags(use_adain=use_adain, train_input=train_input)
    if RefUNet is not None:
        latent = build_pipeline(RefUNet, ref_ap, ucfg, flags, lora_unet)
    else:
        unet = synth.make_unet(ucfg, seed=0, lora_rank=lora_unet)
        orig = synth.make_unet(ucfg, seed=0)
        latent = LatentRestorePipeline(unet, orig, synth.caption_embedding(ucfg.cross_attention_dim), flags)
    vae = synth.make_vae(vcfg, seed=100, lora_rank=lora_vae)
    ovae = synth.make_vae(VaeConfig.tiny(), seed=100)
    if reference_forwards:
        bind_reference_vae_forwards(vae)
        bind_reference_vae_forwards(ovae)
        ovae.decoder.ignore_skip = True
    return ImageRestorePipeline(latent, vae, ovae)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--no-full", action="store_true")
    ap.add_argument("--full-only", action="store_true", help="only the full-width cases (sections 4-7)")
    ap.add_argument("--only", default="", help="comma-separated case names (full-width sections only)")
    args = ap.parse_args()
    only = set(filter(None, args.only.split(",")))
    want = lambda name: not only or name in only
    ref_ap, ref_block, RefUNet = import_reference()
    from oracle import synth
    from oracle.unet import UNetConfig
    GOLDEN.mkdir(parents=True, exist_ok=True)
    torch.manual_seed(0)

    if not (args.full_only or only):
        # 1) the operator: SharedAttnProcessor.forward / AttnProcessor.forward of the reference
        for name, heads, s, n_ref, use_adain, train_input, zeroed in ATTN_CASES:
            attn, hidden, rk, rv = attn_inputs(heads, s, n_ref, zeroed)
            proc = ref_ap.SharedAttnProcessor(self_attn_idx=0 if n_ref else None, save_self_attentions=True,
                                              use_adain=use_adain, train_input=train_input)
            with torch.no_grad():
                out = proc(attn, hidden, ref_keys=[rk] if n_ref else None, ref_values=[rv] if n_ref else None)
            mass = proc.attention_probs.sum(dim=2)  # (B, H, S_k): column mass, used for the per-chunk read-out
            np.savez_compressed(GOLDEN / f"{name}.npz", out=out.numpy(), probs_colsum=mass.numpy(),
                                meta=np.array([heads, s, n_ref, int(use_adain), int(train_input), zeroed]))
            print(name, tuple(out.shape), float(out.abs().max()))
        # the KV-capturing processor of the reference-image UNet
        attn, hidden, _, _ = attn_inputs(2, 64, 0, 0)
        cap_proc = ref_ap.AttnProcessor()
        with torch.no_grad():
            out = cap_proc(attn, hidden)
        np.savez_compressed(GOLDEN / "attn_kv_capture.npz", out=out.numpy(), keys=cap_proc.keys.numpy(), values=cap_proc.values.numpy())

        # the face-embedding cross-attention processor (reference attn_processors.py:100-180)
        out, _ = faceid_case(ref_ap.FaceIDAttnProcessor)
        np.savez_compressed(GOLDEN / "attn_faceid.npz", out=out.numpy())

        # 2) FreeU (reference block.py:3495-3520 on top of the restated fourier_filter)
        g = torch.Generator().manual_seed(11)
        for idx, (h, c) in enumerate([(8, 64), (16, 32)]):
            hs = torch.randn(2, c, h, h, generator=g)
            res = torch.randn(2, c, h, h, generator=g)
            hs2, res2 = ref_block.apply_freeu(idx, hs.clone(), res.clone(), s1=0.9, s2=0.2, b1=1.4, b2=1.6)
            np.savez_compressed(GOLDEN / f"freeu_stage{idx}.npz", hidden=hs2.numpy(), skip=res2.numpy())

        # 3) whole pipeline at the latent boundary on the reduced-width UNet
        tiny = UNetConfig.tiny()
        for name, batch, n_ref, use_adain, train_input, lora_rank, valid in UNET_CASES:
            flags = synth.ModelFlags(use_adain=use_adain, train_input=train_input)
            pipe = build_pipeline(RefUNet, ref_ap, tiny, flags, lora_rank)
            enc, refs, nm, nr = synth.latents(batch, n_ref, tiny.sample_size)
            out = pipe.forward_latents(enc, refs, nm, nr, valid_indices=valid)
            np.savez_compressed(GOLDEN / f"{name}.npz", x0=out.numpy(),
                                meta=np.array([batch, n_ref, int(use_adain), int(train_input), lora_rank]),
                                valid=np.array(valid if valid is not None else [n_ref] * batch))
            print(name, tuple(out.shape), float(out.std()))

        # 3a) the same with the cross-attentions conditioned on face embeddings (FaceIDAttnProcessor, cfg.condition_on_face_embeds;
        #     reference pix2pix_turbo.py:316-320): 2 identities, 3 references, 4 face embeddings each
        faceid_pipeline_case(RefUNet, ref_ap)

        # 3b) VAE alone through the reference's own patched forwards (models/model.py:15-63), and the whole image pipeline
        from oracle.vae import VaeConfig
        for name, use_shortcuts, lora_rank in VAE_CASES:
            vcfg = VaeConfig.tiny()
            vcfg.use_shortcuts = use_shortcuts
            vae = bind_reference_vae_forwards(synth.make_vae(vcfg, seed=100, lora_rank=lora_rank))
            c_t, _, eps_main, _, _, _ = synth.images(2, 1, IMAGE_SIZE, IMAGE_LATENT)
            with torch.no_grad():
                z = vae.encode_sample(c_t, eps_main) * vcfg.scaling_factor
                vae.decoder.incoming_skip_acts = vae.encoder.current_down_blocks
                y = vae.decode(z / vcfg.scaling_factor).clamp(-1, 1)
            np.savez_compressed(GOLDEN / f"{name}.npz", latent=z.numpy(), image=y.numpy().astype(np.float16))
            print(name, tuple(z.shape), float(z.std()), tuple(y.shape), float(y.std()))
        for name, batch, n_ref, use_adain, train_input, lora_unet, lora_vae, use_shortcuts in IMAGE_CASES:
            pipe = tiny_image_models(use_adain, train_input, lora_unet, lora_vae, use_shortcuts, True, RefUNet, ref_ap)
            out = pipe.forward(*synth.images(batch, n_ref, IMAGE_SIZE, IMAGE_LATENT))
            np.savez_compressed(GOLDEN / f"{name}.npz", image=out.numpy().astype(np.float16))
            print(name, tuple(out.shape), float(out.std()))

    # 4) full-width SD-Turbo geometry, released "final model" flags (AdaIN on, refs-only KV), B=1, N=4
    if not args.no_full and want("unet_full_final_n4"):
        full = UNetConfig()
        flags = synth.ModelFlags(use_adain=True, train_input=False)
        pipe = build_pipeline(RefUNet, ref_ap, full, flags, lora_rank=0)
        enc, refs, nm, nr = synth.latents(1, 4, full.sample_size)
        out = pipe.forward_latents(enc, refs, nm, nr)
        np.savez_compressed(GOLDEN / "unet_full_final_n4.npz", x0=out.numpy(), meta=np.array([1, 4, 1, 0, 0]))
        print("unet_full_final_n4", tuple(out.shape), float(out.std()))
    if args.no_full:
        return

    # 5) the shared-attention operator at SD-Turbo widths, reference-count sweep (BASELINE configs[4])
    for name, heads, s, n_ref, use_adain, train_input, zeroed, row_step in ATTN_FULL_CASES:
        if not want(name):
            continue
        attn, hidden, rk, rv = attn_inputs(heads, s, n_ref, zeroed, batch=1)
        proc = ref_ap.SharedAttnProcessor(self_attn_idx=0, save_self_attentions=False, use_adain=use_adain, train_input=train_input)
        with torch.no_grad():
            out = proc(attn, hidden, ref_keys=[rk], ref_values=[rv])
        np.savez_compressed(GOLDEN / f"{name}.npz", out_rows=out[:, ::row_step].numpy(),
                            meta=np.array([heads, s, n_ref, int(use_adain), int(train_input), zeroed, row_step]))
        print(name, tuple(out.shape), float(out.abs().max()), flush=True)

    # 6) latent pipeline at full width: N_ref sweep and the B = 8 batch of BASELINE configs[2] (ragged valid counts)
    full = UNetConfig()
    for name, batch, n_ref, use_adain, train_input, lora_rank, valid in FULL_LATENT_CASES:
        if not want(name):
            continue
        flags = synth.ModelFlags(use_adain=use_adain, train_input=train_input)
        pipe = build_pipeline(RefUNet, ref_ap, full, flags, lora_rank)
        enc, refs, nm, nr = synth.latents(batch, n_ref, full.sample_size)
        out = pipe.forward_latents(enc, refs, nm, nr, valid_indices=valid)
        np.savez_compressed(GOLDEN / f"{name}.npz", x0=out.numpy(),
                            meta=np.array([batch, n_ref, int(use_adain), int(train_input), lora_rank]),
                            valid=np.array(valid if valid is not None else [n_ref] * batch))
        print(name, tuple(out.shape), float(out.std()), flush=True)
        del pipe

    # 7) the BENCHMARKED workload: 512 x 512 images, full UNet + VAE widths, B = 1, N_ref = 4, AdaIN, refs-only KV
    for name, batch, n_ref, use_adain, train_input, lora_unet, lora_vae, use_shortcuts in IMAGE_FULL_CASES:
        if not want(name):
            continue
        pipe = full_image_models(use_adain, train_input, lora_unet, lora_vae, use_shortcuts, True, RefUNet, ref_ap)
        ins = synth.images(batch, n_ref, 512, 64)
        out = pipe.forward(*ins)
        np.savez_compressed(GOLDEN / f"{name}.npz", image=out.numpy().astype(np.float16))
        print(name, tuple(out.shape), float(out.std()), flush=True)


if __name__ == "__main__":
    main()
