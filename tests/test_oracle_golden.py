"""CPU: the oracle restatement against the golden vectors that the REFERENCE'S OWN unet.py / block.py /
attn_processors.py produced (oracle/make_golden.py). fp32 on the CPU, so the bar is bit-exact up to summation order:
we assert max |diff| <= 1e-5 * max |golden|."""
import numpy as np
import pytest
import torch

from oracle import attn_processors as oap
from oracle import synth
from oracle.make_golden import ATTN_CASES, ATTN_FULL_CASES, UNET_CASES, attn_inputs
from oracle.pipeline import LatentRestorePipeline
from oracle.unet import UNetConfig, apply_freeu


def _close(got, want, tol=1e-5):
    want = torch.as_tensor(np.asarray(want))
    scale = float(want.abs().max())
    assert got.shape == want.shape
    assert float((got - want).abs().max()) <= tol * max(scale, 1.0), (float((got - want).abs().max()), scale)


@pytest.mark.parametrize("case", ATTN_CASES, ids=[c[0] for c in ATTN_CASES])
def test_shared_attn_processor_matches_reference(case, golden):
    name, heads, s, n_ref, use_adain, train_input, zeroed = case
    attn, hidden, rk, rv = attn_inputs(heads, s, n_ref, zeroed)
    proc = oap.SharedAttnProcessor(self_attn_idx=0 if n_ref else None, save_self_attentions=True, use_adain=use_adain,
                                   train_input=train_input)
    with torch.no_grad():
        out = proc(attn, hidden, ref_keys=[rk] if n_ref else None, ref_values=[rv] if n_ref else None)
    g = golden(name)
    _close(out, g["out"])
    _close(proc.attention_probs.sum(dim=2), g["probs_colsum"])


@pytest.mark.parametrize("case", ATTN_FULL_CASES, ids=[c[0] for c in ATTN_FULL_CASES])
def test_shared_attn_processor_full_width_matches_reference(case, golden):
    """SD-Turbo widths, reference counts 1 / 2 / 8 (zero-filled padded slots at N = 8) and the 1+4 own-chunk case."""
    name, heads, s, n_ref, use_adain, train_input, zeroed, row_step = case
    attn, hidden, rk, rv = attn_inputs(heads, s, n_ref, zeroed, batch=1)
    proc = oap.SharedAttnProcessor(self_attn_idx=0, use_adain=use_adain, train_input=train_input)
    with torch.no_grad():
        out = proc(attn, hidden, ref_keys=[rk], ref_values=[rv])
    g = golden(name)
    assert list(g["meta"]) == [heads, s, n_ref, int(use_adain), int(train_input), zeroed, row_step]
    _close(out[:, ::row_step], g["out_rows"])


def test_kv_capture_processor_matches_reference(golden):
    attn, hidden, _, _ = attn_inputs(2, 64, 0, 0)
    proc = oap.AttnProcessor()
    with torch.no_grad():
        out = proc(attn, hidden)
    g = golden("attn_kv_capture")
    _close(out, g["out"])
    _close(proc.keys, g["keys"])
    _close(proc.values, g["values"])
    assert proc.is_self_attn is True
    proc.reset()
    assert proc.keys is None and proc.values is None


@pytest.mark.parametrize("idx,h,c", [(0, 8, 64), (1, 16, 32)])
def test_freeu_matches_reference(idx, h, c, golden):
    g = torch.Generator().manual_seed(11)
    cases = []
    for (hh, cc) in [(8, 64), (16, 32)]:
        cases.append((torch.randn(2, cc, hh, hh, generator=g), torch.randn(2, cc, hh, hh, generator=g)))
    hs, res = cases[idx]
    hs2, res2 = apply_freeu(idx, hs.clone(), res.clone(), s1=0.9, s2=0.2, b1=1.4, b2=1.6)
    gold = golden(f"freeu_stage{idx}")
    _close(hs2, gold["hidden"])
    _close(res2, gold["skip"])


def test_freeu_closed_form_equals_fft():
    """fourier_filter(threshold=1, scale=s) == x - (1-s)/(HW) Re(sum of 4 low-frequency terms)."""
    from oracle.diffusers024 import fourier_filter
    g = torch.Generator().manual_seed(3)
    for h, s in [(8, 0.9), (16, 0.2)]:
        x = torch.randn(2, 5, h, h, generator=g, dtype=torch.float64)
        m = torch.arange(h, dtype=torch.float64)
        e = torch.exp(2j * torch.pi * m / h)                       # e^{+2 pi i m/H}: the u = -1 basis
        X00 = x.sum((-1, -2), keepdim=True)
        X10 = (x * e[:, None]).sum((-1, -2), keepdim=True)
        X01 = (x * e[None, :]).sum((-1, -2), keepdim=True)
        X11 = (x * e[:, None] * e[None, :]).sum((-1, -2), keepdim=True)
        corr = (X00 + X10 * e.conj()[:, None] + X01 * e.conj()[None, :] + X11 * e.conj()[:, None] * e.conj()[None, :]).real
        closed = x - (1 - s) / (h * h) * corr
        ref = fourier_filter(x.float(), threshold=1, scale=s)
        assert float((closed.float() - ref).abs().max()) < 5e-6


def _tiny_pipeline(use_adain, train_input, lora_rank):
    tiny = UNetConfig.tiny()
    flags = synth.ModelFlags(use_adain=use_adain, train_input=train_input)
    unet = synth.make_unet(tiny, seed=0, lora_rank=lora_rank)
    orig = synth.make_unet(tiny, seed=0)
    return tiny, LatentRestorePipeline(unet, orig, synth.caption_embedding(tiny.cross_attention_dim), flags)


@pytest.mark.parametrize("case", UNET_CASES, ids=[c[0] for c in UNET_CASES])
def test_tiny_pipeline_matches_reference(case, golden):
    name, batch, n_ref, use_adain, train_input, lora_rank, valid = case
    tiny, pipe = _tiny_pipeline(use_adain, train_input, lora_rank)
    enc, refs, nm, nr = synth.latents(batch, n_ref, tiny.sample_size)
    out = pipe.forward_latents(enc, refs, nm, nr, valid_indices=valid)
    _close(out, golden(name)["x0"])


def test_faceid_pipeline_matches_reference(golden):
    """cfg.condition_on_face_embeds: the cross-attentions become FaceIDAttnProcessors fed with face embeddings
    (reference attn_processors.py:100-180, pix2pix_turbo.py:316-320); golden from the reference's own processors."""
    from oracle.make_golden import FACEID_CASE
    name, batch, n_ref, use_adain, train_input, lora_rank = FACEID_CASE
    tiny = UNetConfig.tiny()
    flags = synth.ModelFlags(use_adain=use_adain, train_input=train_input, condition_on_face_embeds=True)
    pipe = LatentRestorePipeline(synth.make_unet(tiny, seed=0, lora_rank=lora_rank), synth.make_unet(tiny, seed=0),
                                 synth.caption_embedding(tiny.cross_attention_dim), flags)
    synth.seed_face_processors(pipe.unet)
    enc, refs, nm, nr = synth.latents(batch, n_ref, tiny.sample_size)
    out = pipe.forward_latents(enc, refs, nm, nr, face_embeds=synth.face_embeddings(batch))
    _close(out, golden(name)["x0"])
    # the face embeddings matter: other embeddings give a clearly different result
    other = pipe.forward_latents(enc, refs, nm, nr, face_embeds=synth.face_embeddings(batch, seed=6))
    assert float((other - out).abs().max()) > 1e-3


def test_processor_registration_numbering():
    """reference attn_processors.py:282-331: only up_blocks.*.attn1 get self_attn_idx 0..8 in module order."""
    tiny = UNetConfig.tiny()
    unet = synth.make_unet(tiny, seed=0)
    oap.register_attention_processor(unet, synth.ModelFlags(use_adain=True, train_input=False))
    procs = unet.attn_processors
    assert len(procs) == 32
    shared = [(n, p.self_attn_idx) for n, p in procs.items() if p.self_attn_idx is not None]
    assert [i for _, i in shared] == list(range(9))
    assert all(n.startswith("up_blocks") and n.endswith("attn1.processor") for n, _ in shared)
    assert all(p.use_adain and not p.train_input for p in procs.values())
    orig = synth.make_unet(tiny, seed=0)
    oap.register_attention_processor_kv_unet(orig)
    kv = [n for n, p in orig.attn_processors.items() if type(p) is oap.AttnProcessor]
    assert len(kv) == 9 and all("up_blocks" in n and "attn1" in n for n in kv)


def test_padded_slot_takes_softmax_mass(golden):
    """Quirk (pix2pix_turbo.py:269-273): zeroed reference slots are NOT masked — they receive probability mass."""
    g = golden("attn_adain_padded_slot")
    heads, s, n_ref = int(g["meta"][0]), int(g["meta"][1]), int(g["meta"][2])
    mass = torch.as_tensor(g["probs_colsum"])            # (B, H, S_k)
    per_ref = mass.view(2, heads, n_ref, s).sum(-1)
    assert float(per_ref[1, :, -1].min()) > 0.0
    assert abs(float(per_ref.sum(-1).mean()) - s) < 1e-2     # each query row sums to 1


@pytest.mark.slow
def test_full_width_pipeline_matches_reference(golden):
    full = UNetConfig()
    flags = synth.ModelFlags(use_adain=True, train_input=False)
    unet, orig = synth.make_unet(full, seed=0), synth.make_unet(full, seed=0)
    pipe = LatentRestorePipeline(unet, orig, synth.caption_embedding(full.cross_attention_dim), flags)
    enc, refs, nm, nr = synth.latents(1, 4, full.sample_size)
    out = pipe.forward_latents(enc, refs, nm, nr)
    _close(out, golden("unet_full_final_n4")["x0"], tol=1e-4)


# ------------------------------------------------------------------------------------------------ VAE + image pipeline
def _vae_cases():
    from oracle.make_golden import VAE_CASES
    return VAE_CASES


@pytest.mark.parametrize("case", _vae_cases(), ids=[c[0] for c in _vae_cases()])
def test_vae_oracle_matches_reference_forwards(case, golden):
    """oracle/vae.py forwards == the reference's my_vae_encoder_fwd / my_vae_decoder_fwd (models/model.py:15-63)."""
    from oracle.make_golden import IMAGE_LATENT, IMAGE_SIZE
    from oracle.vae import VaeConfig
    name, use_shortcuts, lora_rank = case
    vcfg = VaeConfig.tiny()
    vcfg.use_shortcuts = use_shortcuts
    vae = synth.make_vae(vcfg, seed=100, lora_rank=lora_rank)
    c_t, _, eps_main, _, _, _ = synth.images(2, 1, IMAGE_SIZE, IMAGE_LATENT)
    with torch.no_grad():
        z = vae.encode_sample(c_t, eps_main) * vcfg.scaling_factor
        vae.decoder.incoming_skip_acts = vae.encoder.current_down_blocks
        y = vae.decode(z / vcfg.scaling_factor).clamp(-1, 1)
    g = golden(name)
    _close(z, g["latent"])
    assert float((y - torch.as_tensor(g["image"]).float()).abs().max()) <= 1e-3     # golden images are stored as fp16


def _image_cases():
    from oracle.make_golden import IMAGE_CASES
    return IMAGE_CASES


@pytest.mark.parametrize("case", _image_cases(), ids=[c[0] for c in _image_cases()])
def test_image_pipeline_oracle_matches_reference(case, golden):
    from oracle.make_golden import IMAGE_LATENT, IMAGE_SIZE, tiny_image_models
    name, batch, n_ref, use_adain, train_input, lora_unet, lora_vae, use_shortcuts = case
    pipe = tiny_image_models(use_adain, train_input, lora_unet, lora_vae, use_shortcuts, reference_forwards=False)
    out = pipe.forward(*synth.images(batch, n_ref, IMAGE_SIZE, IMAGE_LATENT))
    assert float((out - torch.as_tensor(golden(name)["image"]).float()).abs().max()) <= 1e-3


def test_faceid_processor_matches_reference(golden):
    from oracle.make_golden import faceid_case
    out, proc = faceid_case(oap.FaceIDAttnProcessor)
    _close(out, golden("attn_faceid")["out"])
    assert proc.is_self_attn is False
