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
