"""GPU parity at the BENCHMARKED geometry (BASELINE.json configs[1..4]): SD-Turbo UNet widths (320/640/1280/1280,
heads 5/10/20/20), sd-vae-ft-mse VAE widths (128/256/512/512), 512 x 512 images, against golden vectors produced by
the reference's own code on the CPU in fp32 (oracle/make_golden.py sections 5-7).

Every case prints and checks two figures: relative L2 over the whole tensor, and max|delta| / max|gold| (the largest
single-element deviation in units of the gold tensor's dynamic range).

Tolerances. north_star asks for 1e-3 relative in fp16. The latent pipeline (two UNet passes) meets rel-L2 <= 1e-3.
One processor call (fp16 input, q/k/v, P, O and output roundings around K = 320..1280 projections) measures
4.5e-4 .. 1.3e-3 on B200 (S = 4096: 4.5-6.7e-4; S = 256 / 1024 at C = 1280 / 640: 0.9-1.3e-3) while the reference
processor under its own precision contract (fp32 module, fp16 autocast, test.py:82-83) on the same inputs and GPU is
1.1e-3 .. 2.1e-3 from the same fp32 gold: the operator does NOT meet 1e-3 at the narrow-S layers, and neither does the
reference. It is held to rel-L2 <= 1.5e-3 AND to <= the reference's own fp16-autocast error measured in the test.
The IMAGE level does NOT meet 1e-3 either: the VAE adds ~60 more fp16-rounded layers and the reference's own
fp16-autocast forward on the same GPU is 2-2.6e-3 away from its fp32 result; the image is held to rel-L2 <= 5e-3 AND
to <= 1.5x the reference's own fp16-autocast error measured in the same test.
"""
import pytest
import torch

from conftest import max_rel, rel_l2

pytestmark = pytest.mark.gpu

OP_TOL = 1e-3
OP_TOL_WIDE = 1.5e-3      # one processor call: measured 4.5e-4 .. 1.32e-3 (reference fp16 autocast: 1.1 .. 2.1e-3)
PIPE_TOL = 1e-3
IMAGE_TOL = 5e-3


@pytest.fixture(scope="module", autouse=True)
def _lib_loaded():
    from instantrestore_b200 import _lib
    assert _lib.load().ir_check_device() == 0


# ------------------------------------------------------------------------------------------------ operator
def _attn_cases():
    from oracle.make_golden import ATTN_FULL_CASES
    return ATTN_FULL_CASES


@pytest.mark.parametrize("case", _attn_cases(), ids=[c[0] for c in _attn_cases()])
def test_shared_attn_full_width_vs_reference_golden(case, golden):
    """SharedAttnProcessor.forward at SD-Turbo widths, reference counts 1 / 2 / 8 (and 1+4 with the own chunk), AdaIN on,
    zero-filled padded slots for N = 8 — the launch shapes of the reference-count sweep (split-KV at B = 1)."""
    from instantrestore_b200.attn_processors import SharedAttnProcessor
    from oracle.make_golden import attn_inputs
    name, heads, s, n_ref, use_adain, train_input, zeroed, row_step = case
    attn, hidden, rk, rv = attn_inputs(heads, s, n_ref, zeroed, batch=1)
    proc = SharedAttnProcessor(self_attn_idx=0, use_adain=use_adain, train_input=train_input)
    out = proc(attn.cuda(), hidden.cuda().half(), ref_keys=[rk.cuda().half()], ref_values=[rv.cuda().half()])
    gold = torch.as_tensor(golden(name)["out_rows"])
    got = out[:, ::row_step].float()
    assert got.shape == gold.shape
    e2, em = rel_l2(got, gold), max_rel(got, gold)
    # the reference's own precision contract on the same inputs and GPU (fp32 module under fp16 autocast, fp16 K/V
    # from the autocast reference UNet): what "fp16 tolerance" means for this operator at this width
    from oracle import attn_processors as oap
    ref_proc = oap.SharedAttnProcessor(self_attn_idx=0, use_adain=use_adain, train_input=train_input)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        ac = ref_proc(attn.cuda(), hidden.cuda(), ref_keys=[rk.cuda().half()], ref_values=[rv.cuda().half()])
    a2, am = rel_l2(ac[:, ::row_step].float(), gold), max_rel(ac[:, ::row_step].float(), gold)
    print(f"{name}: rel-L2 {e2:.3e}  max|d|/max|gold| {em:.3e} | reference fp16-autocast rel-L2 {a2:.3e} max {am:.3e}")
    assert e2 <= OP_TOL_WIDE
    assert e2 <= a2            # never worse than the reference's own fp16 contract
    assert em <= 4e-3


# ------------------------------------------------------------------------------------------------ latent pipeline
def _engine(use_adain, train_input, lora_rank, use_cuda_graph=True):
    from instantrestore_b200.pipeline import ModelFlags, RestoreEngine
    from instantrestore_b200.unet_engine import UNetSpec
    from oracle import synth
    from oracle.unet import UNetConfig
    cfg = UNetConfig()
    unet, orig = synth.make_unet(cfg, seed=0, lora_rank=lora_rank), synth.make_unet(cfg, seed=0)
    spec = UNetSpec(block_out_channels=tuple(cfg.block_out_channels), attention_head_dim=tuple(cfg.attention_head_dim),
                    cross_attention_dim=cfg.cross_attention_dim)
    return RestoreEngine(unet.state_dict(), orig.state_dict(), synth.caption_embedding(cfg.cross_attention_dim),
                         ModelFlags(use_adain=use_adain, train_input=train_input), spec=spec, use_cuda_graph=use_cuda_graph)


def _autocast_error(use_adain, train_input, lora_rank, batch, n_ref, valid, gold):
    """The reference's own precision contract (fp32 weights + fp16 autocast, test.py:82-83) on this GPU vs the fp32 gold."""
    from oracle import synth
    from oracle.pipeline import LatentRestorePipeline
    from oracle.unet import UNetConfig
    cfg = UNetConfig()
    flags = synth.ModelFlags(use_adain=use_adain, train_input=train_input)
    pipe = LatentRestorePipeline(synth.make_unet(cfg, seed=0, lora_rank=lora_rank).cuda(), synth.make_unet(cfg, seed=0).cuda(),
                                 synth.caption_embedding(cfg.cross_attention_dim).cuda(), flags)
    enc, refs, nm, nr = (t.cuda() for t in synth.latents(batch, n_ref, cfg.sample_size))
    outs = []
    with torch.autocast("cuda", dtype=torch.float16):
        for b in range(batch):     # one identity at a time: the eager path materialises (B*H, S, S_k) score tensors
            v = None if valid is None else [valid[b]]
            outs.append(pipe.forward_latents(enc[b:b + 1], refs[b:b + 1], nm[b:b + 1], nr[b * n_ref:(b + 1) * n_ref], valid_indices=v))
    out = torch.cat(outs).float()
    del pipe
    torch.cuda.empty_cache()
    return rel_l2(out, gold)


def _latent_cases():
    from oracle.make_golden import FULL_LATENT_CASES
    return FULL_LATENT_CASES


@pytest.mark.parametrize("case", _latent_cases(), ids=[c[0] for c in _latent_cases()])
def test_full_width_latent_pipeline_vs_reference_golden(case, golden):
    """Reference-count sweep N_ref in {1, 2, 8} (configs[4]) and the B = 8 batch with ragged valid counts (configs[2]):
    different tile / split-K / CTA-pair / split-KV policies than B = 1, N = 4."""
    from oracle import synth
    name, batch, n_ref, use_adain, train_input, lora_rank, valid = case
    eng = _engine(use_adain, train_input, lora_rank)
    enc, refs, nm, nr = (t.cuda() for t in synth.latents(batch, n_ref, 64))
    gold = torch.as_tensor(golden(name)["x0"])
    out = eng.forward_latents(enc, refs, nm, nr, valid_indices=valid).float().cpu()
    del eng
    torch.cuda.empty_cache()
    e2, em = rel_l2(out, gold), max_rel(out, gold)
    ac = _autocast_error(use_adain, train_input, lora_rank, batch, n_ref, valid, gold)
    print(f"{name}: rel-L2 {e2:.3e}  max|d|/max|gold| {em:.3e}  reference-autocast rel-L2 {ac:.3e}")
    assert e2 <= PIPE_TOL
    assert e2 <= 1.5 * ac + 2e-4
    assert em <= 1e-2
    if batch > 1:   # per-identity figures: no identity of the batch hides behind the others
        worst = max(rel_l2(out[b], gold[b]) for b in range(batch))
        print(f"{name}: worst identity rel-L2 {worst:.3e}")
        assert worst <= 1.5e-3


def test_batch8_rows_equal_single_identity_results():
    """configs[2] (B = 8 on one GPU) vs configs[1] (B = 1): every row of the B = 8 step equals the B = 1 result of the same
    identity within fp16 tolerance (different kernels are chosen, so not bit-identical)."""
    from oracle import synth
    eng = _engine(True, False, 4)
    enc, refs, nm, nr = (t.cuda() for t in synth.latents(8, 4, 64))
    full = eng.forward_latents(enc, refs, nm, nr).clone()
    for b in range(8):
        one = eng.forward_latents(enc[b:b + 1].contiguous(), refs[b:b + 1].contiguous(), nm[b:b + 1].contiguous(),
                                  nr[4 * b:4 * b + 4].contiguous())
        e2, em = rel_l2(one[0], full[b]), max_rel(one[0], full[b])
        print(f"identity {b}: B=8 row vs B=1 result rel-L2 {e2:.3e} max {em:.3e}")
        assert e2 <= 1e-3


# ------------------------------------------------------------------------------------------------ image pipeline (the bench workload)
def _full_pipeline(use_adain, train_input, lora_unet, lora_vae, use_shortcuts, graph=True):
    from instantrestore_b200.pipeline import ModelFlags, RestorePipeline
    from instantrestore_b200.unet_engine import UNetSpec
    from oracle import synth
    from oracle.unet import UNetConfig
    from oracle.vae import VaeConfig
    ucfg, vcfg = UNetConfig(), VaeConfig(use_shortcuts=use_shortcuts)
    unet, orig = synth.make_unet(ucfg, seed=0, lora_rank=lora_unet), synth.make_unet(ucfg, seed=0)
    vae, ovae = synth.make_vae(vcfg, seed=100, lora_rank=lora_vae), synth.make_vae(VaeConfig(), seed=100)
    spec = UNetSpec(block_out_channels=tuple(ucfg.block_out_channels), attention_head_dim=tuple(ucfg.attention_head_dim),
                    cross_attention_dim=ucfg.cross_attention_dim)
    return RestorePipeline(unet.state_dict(), orig.state_dict(), vae.state_dict(), ovae.state_dict(),
                           synth.caption_embedding(ucfg.cross_attention_dim), ModelFlags(use_adain=use_adain, train_input=train_input),
                           spec=spec, vae_block_out_channels=vcfg.block_out_channels, use_shortcuts=use_shortcuts, use_cuda_graph=graph)


def _image_cases():
    from oracle.make_golden import IMAGE_FULL_CASES
    return IMAGE_FULL_CASES


@pytest.mark.parametrize("case", _image_cases(), ids=[c[0] for c in _image_cases()])
def test_benchmarked_image_pipeline_vs_reference_golden(case, golden):
    """THE bench.py workload: 512 x 512 degraded image + 4 references in, restored 512 x 512 image out, full UNet and VAE
    widths, AdaIN, refs-only KV, through the CUDA graph — against the reference code's fp32 CPU result."""
    from oracle import synth
    from oracle.make_golden import full_image_models
    name, batch, n_ref, use_adain, train_input, lora_unet, lora_vae, use_shortcuts = case
    pipe = _full_pipeline(use_adain, train_input, lora_unet, lora_vae, use_shortcuts)
    c_t, cond, eps_main, eps_ref, noise_main, noise_ref = synth.images(batch, n_ref, 512, 64)
    out, _, _ = pipe.forward(c_t.cuda().half(), conditioning_images=cond.cuda().half(), valid_indices=[n_ref] * batch,
                             eps_main=eps_main, eps_ref=eps_ref, noise_main=noise_main, noise_ref=noise_ref)
    out = out.float().cpu()
    assert out.shape == c_t.shape
    out2, _, _ = pipe.forward(c_t.cuda().half(), conditioning_images=cond.cuda().half(), eps_main=eps_main, eps_ref=eps_ref,
                              noise_main=noise_main, noise_ref=noise_ref)
    assert torch.equal(out2.float().cpu(), out)                 # graph replay is deterministic
    del pipe
    torch.cuda.empty_cache()
    gold = torch.as_tensor(golden(name)["image"]).float()
    e2, em = rel_l2(out, gold), max_rel(out, gold)
    ref = full_image_models(use_adain, train_input, lora_unet, lora_vae, use_shortcuts, reference_forwards=False)
    for m in (ref.latent.unet, ref.latent.original_unet, ref.vae, ref.original_vae):
        m.cuda()
    ref.latent.caption_enc = ref.latent.caption_enc.cuda()
    with torch.autocast("cuda", dtype=torch.float16):
        ac = ref.forward(c_t.cuda(), cond.cuda(), eps_main.cuda(), eps_ref.cuda(), noise_main.cuda(), noise_ref.cuda()).float().cpu()
    a2, am = rel_l2(ac, gold), max_rel(ac, gold)
    print(f"{name}: ours rel-L2 {e2:.3e} max|d|/max|gold| {em:.3e} | reference fp16-autocast rel-L2 {a2:.3e} max {am:.3e}")
    assert e2 <= IMAGE_TOL
    assert e2 <= 1.5 * a2 + 3e-4
    assert em <= 1.5 * am + 2e-2


def test_processors_drive_the_oracle_unet_with_lora():
    """Drop-in path with peft-wrapped attention projections (reference pix2pix_turbo.py:171-179): the processors must
    apply W + scaling * B @ A, not the base weight alone."""
    from instantrestore_b200 import attn_processors as ours
    from oracle import synth
    from oracle.pipeline import LatentRestorePipeline
    from oracle.unet import UNetConfig
    tiny = UNetConfig.tiny()
    flags = synth.ModelFlags(use_adain=True, train_input=False)
    mk = lambda: (synth.make_unet(tiny, seed=0, lora_rank=4, lora_b_std=0.05), synth.make_unet(tiny, seed=0))
    unet, orig = mk()
    pipe = LatentRestorePipeline(unet.cuda(), orig.cuda(), synth.caption_embedding(tiny.cross_attention_dim).cuda(), flags, processors=ours)
    enc, refs, nm, nr = (t.cuda() for t in synth.latents(2, 2, tiny.sample_size))
    with torch.autocast("cuda", dtype=torch.float16):
        out = pipe.forward_latents(enc, refs, nm, nr)
    unet_c, orig_c = mk()
    gold = LatentRestorePipeline(unet_c, orig_c, synth.caption_embedding(tiny.cross_attention_dim), flags).forward_latents(
        *synth.latents(2, 2, tiny.sample_size))
    # the LoRA delta matters at this scale: dropping it must be far outside the tolerance
    base_u, base_o = synth.make_unet(tiny, seed=0, lora_rank=0), synth.make_unet(tiny, seed=0)
    no_lora = LatentRestorePipeline(base_u, base_o, synth.caption_embedding(tiny.cross_attention_dim), flags).forward_latents(
        *synth.latents(2, 2, tiny.sample_size))
    # the same fp32 oracle UNet under autocast with the REFERENCE processors: the error floor of this harness (the
    # convolutions / norms around the processors run in torch fp16 autocast in both arms)
    unet_a, orig_a = mk()
    pipe_a = LatentRestorePipeline(unet_a.cuda(), orig_a.cuda(), synth.caption_embedding(tiny.cross_attention_dim).cuda(), flags)
    with torch.autocast("cuda", dtype=torch.float16):
        out_a = pipe_a.forward_latents(enc, refs, nm, nr)
    e2, a2 = rel_l2(out.float().cpu(), gold), rel_l2(out_a.float().cpu(), gold)
    print(f"drop-in with LoRA: rel-L2 {e2:.3e}; reference processors under the same autocast {a2:.3e} "
          f"(dropping the LoRA delta would be {rel_l2(no_lora, gold):.3e})")
    assert rel_l2(no_lora, gold) > 10 * PIPE_TOL
    assert e2 <= 1.5e-3
    assert e2 <= 1.25 * a2 + 1e-4
