"""CPU: host-side logic of the product package — checkpoint key layout, LoRA merge, weight packing, scheduler
constants, time-embedding fold — checked against the oracle."""
import math

import pytest
import torch

from instantrestore_b200.pipeline import ddpm_coeffs
from instantrestore_b200.synthetic import synthetic_unet_state_dict, unet_parameter_shapes
from instantrestore_b200.unet_engine import UNetSpec, timestep_embedding
from instantrestore_b200.weights import StateDictView, conv_weight_khwc, geglu_interleave_index
from oracle import synth
from oracle.diffusers024 import DDPMScheduler1Step, LoraConv2d, LoraLinear, Timesteps
from oracle.unet import UNet2DConditionModel, UNetConfig


def _tiny_spec():
    t = UNetConfig.tiny()
    return t, UNetSpec(block_out_channels=t.block_out_channels, attention_head_dim=t.attention_head_dim,
                       cross_attention_dim=t.cross_attention_dim)


@pytest.mark.parametrize("lora_rank", [0, 4])
def test_synthetic_checkpoint_has_reference_key_layout(lora_rank):
    """The product's synthetic checkpoint loads strict=True into the reference-shaped module tree."""
    cfg, spec = _tiny_spec()
    sd = synthetic_unet_state_dict(spec, seed=3, lora_rank=lora_rank)
    model = UNet2DConditionModel(cfg)
    if lora_rank:
        from oracle.diffusers024 import add_lora
        add_lora(model, synth.UNET_LORA_TARGETS, r=lora_rank, alpha=lora_rank // 2)
    model.load_state_dict(sd, strict=True)
    if lora_rank:
        assert "down_blocks.0.resnets.0.conv1.base_layer.weight" in sd
        assert "down_blocks.0.resnets.0.conv1.lora_A.default.weight" in sd
        assert "conv_in.weight" in sd and "conv_in.base_layer.weight" not in sd          # pix2pix_turbo.py:205
        assert "time_embedding.linear_1.weight" in sd
        assert "up_blocks.1.attentions.0.transformer_blocks.0.attn1.to_out.0.lora_B.default.weight" in sd


def test_full_size_shapes_match_survey_census():
    shapes = unet_parameter_shapes(UNetSpec())
    convs = [s for s in shapes if s[2] == "conv"]
    assert len(convs) == 66                                    # 66 convolutions
    n_params = sum(math.prod(s[1]) for s in shapes if s[2] != "norm")
    assert 860e6 < n_params < 870e6                            # SD-2.1 UNet: ~866 M parameters


def test_lora_merge_equals_unmerged_forward():
    """W' = W + (alpha/r) B A reproduces base(x) + B(A(x)) * alpha/r for Linear and Conv2d (peft 0.10 semantics)."""
    g = torch.Generator().manual_seed(0)
    lin = LoraLinear(torch.nn.Linear(16, 24), r=4, alpha=2)
    conv = LoraConv2d(torch.nn.Conv2d(8, 12, 3, padding=1), r=4, alpha=2)
    for m in (lin, conv):
        torch.nn.init.normal_(m.lora_B["default"].weight, std=0.3, generator=g)
    sd = {f"lin.{k}": v for k, v in lin.state_dict().items()}
    sd.update({f"conv.{k}": v for k, v in conv.state_dict().items()})
    v = StateDictView(sd)
    x = torch.randn(5, 16, generator=g)
    assert torch.allclose(x @ v.weight("lin").T + v.bias("lin"), lin(x), atol=1e-5)
    xi = torch.randn(2, 8, 6, 6, generator=g)
    merged = torch.nn.functional.conv2d(xi, v.weight("conv"), v.bias("conv"), padding=1)
    assert torch.allclose(merged, conv(xi), atol=1e-5)
    assert v.has("lin.weight") and v.has("conv.weight") and not v.has("nope.weight")


def test_lora_adapter_name_vae_skip_is_merged():
    lin = LoraLinear(torch.nn.Linear(8, 8), r=2, alpha=1, adapter="vae_skip")
    torch.nn.init.normal_(lin.lora_B["vae_skip"].weight, std=0.3)
    v = StateDictView({f"m.{k}": t for k, t in lin.state_dict().items()})
    x = torch.randn(3, 8)
    assert torch.allclose(x @ v.weight("m").T + v.bias("m"), lin(x), atol=1e-5)


def test_conv_weight_packing_is_tap_major():
    w = torch.randn(6, 5, 3, 3)
    p = conv_weight_khwc(w, c_in_pad=8).float().view(6, 3, 3, 8)
    assert torch.allclose(p[..., :5], w.permute(0, 2, 3, 1).half().float())
    assert float(p[..., 5:].abs().max()) == 0.0


def test_geglu_interleave():
    idx = geglu_interleave_index(512)
    assert sorted(idx.tolist()) == list(range(512))
    assert idx[:64].tolist() == list(range(64)) and idx[64:128].tolist() == list(range(256, 320))


def test_scheduler_constants_match_oracle():
    sch = DDPMScheduler1Step()
    for t in (1, 249, 999):
        a, s = ddpm_coeffs(t)
        oa, os_ = sch.coeffs(t)
        assert abs(a - oa) < 1e-7 and abs(s - os_) < 1e-7
        assert abs(a * a + s * s - 1.0) < 1e-6


def test_timestep_embedding_matches_oracle():
    tp = Timesteps(320, True, 0)
    for t in (1, 249):
        assert torch.allclose(timestep_embedding(t, 320), tp(torch.tensor([t])), atol=1e-6)


def test_shard_ranges_partition_identities():
    from instantrestore_b200.dist import shard_range
    for n, world in [(64, 8), (32, 8), (7, 3), (1, 2), (0, 4)]:
        seen = []
        for r in range(world):
            lo, hi = shard_range(n, r, world)
            seen += list(range(lo, hi))
        assert seen == list(range(n))
        sizes = [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]
        assert max(sizes) - min(sizes) <= 1


@pytest.mark.parametrize("use_shortcuts,lora_rank", [(False, 0), (True, 4)])
def test_synthetic_vae_checkpoint_has_reference_key_layout(use_shortcuts, lora_rank):
    from instantrestore_b200.synthetic import synthetic_vae_state_dict
    from oracle.diffusers024 import add_lora
    from oracle.vae import VAE_LORA_TARGETS, AutoencoderKL, VaeConfig
    cfg = VaeConfig.tiny()
    cfg.use_shortcuts = use_shortcuts
    m = AutoencoderKL(cfg)
    if lora_rank:
        extra = ["skip_conv_1", "skip_conv_2", "skip_conv_3", "skip_conv_4"] if use_shortcuts else []
        add_lora(m, list(VAE_LORA_TARGETS) + extra, r=lora_rank, alpha=lora_rank // 2, adapter="vae_skip")
    m.load_state_dict(synthetic_vae_state_dict(cfg.block_out_channels, lora_rank=lora_rank, use_shortcuts=use_shortcuts), strict=True)


def test_full_size_vae_parameter_count():
    from instantrestore_b200.synthetic import vae_parameter_shapes
    n = sum(math.prod(s[1]) * (1 if s[2] == "conv_nobias" else 1) + (s[1][0] if s[2] in ("conv", "linear", "norm") else 0)
            for s in vae_parameter_shapes())
    assert 83.5e6 < n < 83.8e6                                  # sd-vae-ft-mse: 83.65 M parameters


def test_checkpoint_splitting_and_cfg_decoding():
    from instantrestore_b200.inference import decode_cfg, split_state_dict
    sd = {"net.unet.conv_in.weight": 1, "net.module.vae.encoder.conv_in.weight": 2, "net.original_unet.a.b": 3,
          "net.original_vae.quant_conv.bias": 4, "net.text_encoder.x": 5}
    parts = split_state_dict(sd)
    assert set(parts) == {"unet", "vae", "original_unet", "original_vae", "text_encoder"}
    assert parts["vae"] == {"encoder.conv_in.weight": 2} and parts["unet"] == {"conv_in.weight": 1}
    cfg = decode_cfg({"model": {"use_adain": True, "train_input": False, "lora_rank_unet": 32}, "data": {"max_conditioning_images": 4, "x": 1}})
    assert cfg.model.use_adain and not cfg.model.train_input and cfg.model.use_shared_attention
    assert cfg.data.max_conditioning_images == 4
    assert decode_cfg(None).model.train_input is True          # ModelConfig defaults (train_config.py:118-147)


def test_image_transform_matches_reference_transform():
    """test.py:54-59: Resize(512, LANCZOS) + CenterCrop(512) + ToTensor + Normalize(0.5, 0.5)."""
    import numpy as np
    from PIL import Image
    from torchvision import transforms
    from instantrestore_b200.inference import image_to_tensor, tensor2im
    tf = transforms.Compose([transforms.Resize(512, interpolation=transforms.InterpolationMode.LANCZOS),
                             transforms.CenterCrop(512), transforms.ToTensor(),
                             transforms.Normalize([0.5, 0.5, 0.5], [0.5, 0.5, 0.5])])
    rng = np.random.default_rng(0)
    for (h, w) in [(512, 512), (300, 400), (700, 600), (1024, 512)]:
        im = Image.fromarray((rng.random((h, w, 3)) * 255).astype("uint8"))
        assert torch.equal(image_to_tensor(im), tf(im))
    x = torch.rand(3, 8, 8) * 2.4 - 1.2
    ref = ((x * 0.5 + 0.5).clamp(0, 1) * 255).permute(1, 2, 0).numpy().astype("uint8")     # vis_utils.py:14-23
    assert np.array_equal(np.asarray(tensor2im(x, unnorm=True)), ref)


def test_fused_groupnorm_statistics_shape_rules():
    """Which conv outputs can carry GroupNorm pass A out of the epilogue (ir_conv_gemm.gn_partial): whole groups per
    32-column accumulator chunk (4, 8 or 16 channels per group) and whole 32-pixel slabs per image (hw % 128 == 0)."""
    from instantrestore_b200 import _lib as L
    for hw, c, ok in [(512 * 512, 128, True), (256 * 256, 256, True), (64 * 64, 512, True), (64 * 64, 320, False),
                      (32 * 32, 640, False), (8 * 8, 512, False), (16 * 8, 128, True), (4096, 64, False), (4096, 1024, False)]:
        assert L.gn_partial_supported(hw, c) is ok, (hw, c)
    assert L.gn_partial_numel(3, 4096) == 3 * 128 * 32 * 2
    # the VAE of the step: every GroupNorm input at 512x512 images qualifies
    for hw, c in [(512 * 512, 128), (256 * 256, 128), (256 * 256, 256), (128 * 128, 256), (128 * 128, 512), (64 * 64, 512)]:
        assert L.gn_partial_supported(hw, c)


def test_single_launch_groupnorm_plan_rules():
    """ir_groupnorm_fused_supported is host logic: tensors up to 12 MB whose rows split into whole groups x whole
    16-byte vectors take the one-launch cluster kernel; the big VAE tensors keep the streaming path."""
    import ctypes
    from instantrestore_b200 import _lib
    lib = _lib.load()
    ok = lambda b, hw, c, g=32: bool(lib.ir_groupnorm_fused_supported(b, hw, c, g))
    for shape in [(1, 4096, 320), (4, 4096, 320), (1, 4096, 960), (1, 1024, 1920), (1, 256, 2560), (1, 64, 1280), (2, 4096, 512),
                  (1, 16384, 256), (3, 256, 64), (2, 48, 320)]:
        assert ok(*shape), shape
    for shape in [(1, 512 * 512, 128), (4, 256 * 256, 256), (32, 4096, 320), (5, 4096, 512), (1, 4097, 320)]:
        assert not ok(*shape), shape
    assert not ok(1, 4096, 324)            # channels % groups != 0
    assert not ok(0, 4096, 320)


def test_new_entry_points_reject_null_arguments_without_a_gpu():
    import ctypes
    from instantrestore_b200 import _lib
    lib = _lib.load()
    assert lib.ir_resample_u8_pass(None, 3, 3, 1, 1, None, None, 1, 0, None, 1, 1, 1, 0, None) == -5
    assert lib.ir_u8_to_f16(None, 3, 3, 1, 1, None, None) == -5
    assert lib.ir_image_out_u8(None, None, 1, 1, None) == -5
    assert lib.ir_set_pdl(0) in (0, 1)
    assert lib.ir_set_pdl(0) == 0


def test_upsample_conv_weight_fold_matches_upsample_then_conv():
    """weights.upsample_conv_weight: the four 2x2 sub-pixel kernels of nearest-2x + conv3x3 (diffusers Upsample2D; reference
    block.py:2366,2476), applied on the low-resolution tensor with the tap offsets ir_conv_gemm uses, reproduce
    F.interpolate + F.conv2d (fp64: the fold itself is exact)."""
    import torch
    import torch.nn.functional as F
    from instantrestore_b200.weights import upsample_conv_weight
    g = torch.Generator().manual_seed(0)
    B, Ci, Co, H, W = 2, 3, 5, 4, 6
    x = torch.randn(B, Ci, H, W, generator=g, dtype=torch.float64)
    w = torch.randn(Co, Ci, 3, 3, generator=g, dtype=torch.float64)
    ref = F.conv2d(F.interpolate(x, scale_factor=2.0, mode="nearest"), w, padding=1)
    folded = upsample_conv_weight(w)
    assert folded.dtype == torch.float16 and folded.shape == (4 * Co, 4 * Ci)
    # exact fold in fp64 (the product path rounds it to fp16): redo the sums here
    wk = w.permute(0, 2, 3, 1)
    rows = {0: ((0,), (1, 2)), 1: ((0, 1), (2,))}
    out = torch.zeros(B, Co, 2 * H, 2 * W, dtype=torch.float64)
    xp = F.pad(x, (1, 1, 1, 1))                                   # zero border == TMA out-of-bounds fill
    for py in (0, 1):
        for px in (0, 1):
            acc = torch.zeros(B, Co, H, W, dtype=torch.float64)
            for a, kys in enumerate(rows[py]):
                for b, kxs in enumerate(rows[px]):
                    tap = sum(wk[:, ky, kx] for ky in kys for kx in kxs)               # [Co, Ci]
                    dy, dx = py - 1 + a, px - 1 + b
                    patch = xp[:, :, 1 + dy:1 + dy + H, 1 + dx:1 + dx + W]
                    acc += torch.einsum("oc,bchw->bohw", tap, patch)
                    got = folded.view(4, Co, 2, 2, Ci)[py * 2 + px, :, a, b].double()
                    assert torch.allclose(got, tap, atol=2e-3, rtol=2e-3)                # fp16 rounding of the fold
            out[:, :, py::2, px::2] = acc
    assert torch.allclose(out, ref, atol=1e-12)
