"""CPU checks of the reference-facing surface: the `face_replace.*` import paths of the reference resolve to this
repo's implementation, configuration decoding accepts the reference's YAML / checkpoint `cfg` layout, the LoRA merge
the drop-in processors apply equals what a peft-wrapped projection computes, and model geometry is read from a
checkpoint's state_dict. No kernels run here."""
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]


def test_reference_import_paths_resolve_to_this_repo():
    from face_replace.inference.test import Predictor, run_folder
    from face_replace.models.attn_processors import (AttnProcessor, FaceIDAttnProcessor, SharedAttnProcessor, adain,
                                                     register_attention_processor, register_attention_processor_kv_unet)
    from face_replace.models.face_replace_model import FaceReplaceModel
    from face_replace.models.pix2pix_turbo import Pix2Pix_Turbo
    from face_replace.training.utils.vis_utils import tensor2im
    import face_replace
    import instantrestore_b200.attn_processors as ours
    import instantrestore_b200.inference as inf
    assert Path(face_replace.__file__).resolve().parent == ROOT / "face_replace"
    assert Predictor is inf.Predictor and FaceReplaceModel is inf.FaceReplaceModel
    assert SharedAttnProcessor is ours.SharedAttnProcessor and register_attention_processor is ours.register_attention_processor
    assert callable(run_folder) and callable(tensor2im) and callable(adain)
    assert Pix2Pix_Turbo.__name__ == "RestorePipeline"
    # constructor signatures of the reference (attn_processors.py:186, :102)
    p = SharedAttnProcessor(self_attn_idx=3, save_self_attentions=True, use_adain=True, train_input=False)
    assert (p.self_attn_idx, p.save_self_attentions, p.use_adain, p.train_input, p.attention_probs) == (3, True, True, False, None)
    k = AttnProcessor()
    assert k.keys is None and k.values is None and k.is_self_attn is None
    f = FaceIDAttnProcessor(hidden_size=128, cross_attention_dim=256, embed_dim=512)
    assert sorted(f.state_dict()) == ["face_projection.bias", "face_projection.weight", "to_k_face_embed.weight", "to_v_face_embed.weight"]
    assert callable(register_attention_processor_kv_unet)


def test_train_config_decodes_reference_yaml_layout(tmp_path):
    from face_replace.configs.train_config import ModelConfig, TrainConfig, decode
    # defaults of the reference ModelConfig / DataConfig (configs/train_config.py:111,118-147)
    d = TrainConfig.from_dict(None)
    assert (d.model.net_type, d.model.lora_rank_unet, d.model.use_adain, d.model.train_input, d.model.noise_timestep) == \
        ("pix2pix_turbo", 16, False, True, 249)
    assert d.data.max_conditioning_images == 4
    # the released final-model file (config_files/train_landmarkloss_adain.yaml), reproduced here as text
    y = tmp_path / "final.yaml"
    y.write_text("compute:\n  batch_size: 1\noptim:\n  lambda_landmark: 5000.0\n  scheduler_type: CONSTANT\n"
                 "data:\n  data_root: path/to/your/data\n  max_conditioning_images: 4\n  store_landmarks: false\n"
                 "model:\n  net_type: pix2pix_turbo\n  lora_rank_unet: 32\n  lora_rank_vae: 32\n  use_shared_attention: true\n"
                 "  guidance_scale: 0\n  use_shortcuts: false\n  use_adain: true\n  train_input: false\n  checkpoint_path: null\n"
                 "  some_future_flag: 7\nlog:\n  vis_attention: false\nsteps:\n  max_steps: 50000\n")
    c = TrainConfig.from_yaml(y)
    assert isinstance(c.model, ModelConfig)
    assert (c.model.lora_rank_unet, c.model.use_adain, c.model.train_input, c.model.use_shortcuts) == (32, True, False, False)
    assert c.model.some_future_flag == 7 and c.optim.lambda_landmark == 5000.0 and c.steps.max_steps == 50000
    assert decode(TrainConfig, c) is c
    assert decode(TrainConfig, {"model": {"use_adain": True}}).model.use_adain is True
    ref_yaml = Path("/root/reference/config_files/train_base.yaml")
    if ref_yaml.exists():   # build container only
        b = TrainConfig.from_yaml(ref_yaml)
        assert b.model.use_adain is False and b.model.lora_rank_unet == 32


def test_natural_sort_of_conditioning_files():
    from face_replace.inference.test import _natural_key
    names = [Path(n) for n in ("10.png", "2.png", "1.png", "ref_12.png", "ref_3.png")]
    assert [p.name for p in sorted(names, key=_natural_key)] == ["1.png", "2.png", "10.png", "ref_3.png", "ref_12.png"]


def test_effective_weight_merges_lora_like_the_wrapped_module():
    """peft's `.weight` is the base weight only; the processors must run W + scaling * B @ A."""
    from instantrestore_b200.attn_processors import _effective_weight
    from oracle.diffusers024 import LoraLinear
    g = torch.Generator().manual_seed(3)
    base = torch.nn.Linear(48, 32, bias=True)
    mod = LoraLinear(base, r=4, alpha=2.0)
    with torch.no_grad():
        mod.lora_A["default"].weight.copy_(torch.randn(4, 48, generator=g))
        mod.lora_B["default"].weight.copy_(torch.randn(32, 4, generator=g))
    w, params = _effective_weight(mod)
    x = torch.randn(5, 48, generator=g)
    assert torch.allclose(x @ w.T + base.bias, mod(x), atol=1e-5)
    assert not torch.allclose(w, base.weight)                 # the delta is really there
    assert len(params) == 3
    # peft-style dict scaling + several adapters, one inactive
    mod.scaling = {"default": 0.5, "other": 3.0}
    mod.lora_A["other"] = torch.nn.Linear(48, 4, bias=False)
    mod.lora_B["other"] = torch.nn.Linear(4, 32, bias=False)
    mod.active_adapters = ["default"]
    w2, _ = _effective_weight(mod)
    want = base.weight + 0.5 * mod.lora_B["default"].weight @ mod.lora_A["default"].weight
    assert torch.allclose(w2, want, atol=1e-6)
    mod.merged = True
    assert torch.equal(_effective_weight(mod)[0], base.weight.float())
    plain, p = _effective_weight(torch.nn.Linear(8, 8))
    assert plain.shape == (8, 8) and len(p) == 1


def test_proj_cache_key_tracks_lora_parameters():
    from instantrestore_b200.attn_processors import _ProjCache
    from oracle.diffusers024 import Attention, add_lora
    attn = Attention(query_dim=64, heads=1, dim_head=64)
    add_lora(attn, ["to_q", "to_k", "to_v", "to_out.0"], r=2, alpha=1, b_std=0.1, generator=torch.Generator().manual_seed(0))
    cache = _ProjCache()
    w1 = cache.get(attn, True)
    assert w1["qkv"].shape == (192, 64) and w1["qkvb"] is None and w1["ob"] is not None
    assert cache.get(attn, True) is w1                                      # unchanged parameters: cached
    with torch.no_grad():
        attn.to_q.lora_B["default"].weight.add_(1.0)
    w2 = cache.get(attn, True)
    assert w2 is not w1 and not torch.equal(w2["q"], w1["q"]) and torch.equal(w2["k"], w1["k"])


def test_geometry_is_read_from_the_state_dict():
    from instantrestore_b200.synthetic import synthetic_unet_state_dict, synthetic_vae_state_dict
    from instantrestore_b200.unet_engine import UNetSpec
    from instantrestore_b200.weights import infer_unet_geometry, infer_vae_channels
    spec = UNetSpec(block_out_channels=(64, 128, 256, 256), attention_head_dim=(1, 2, 4, 4), cross_attention_dim=128)
    for rank in (0, 4):
        g = infer_unet_geometry(synthetic_unet_state_dict(spec, seed=0, lora_rank=rank))
        assert g["block_out_channels"] == (64, 128, 256, 256) and g["attention_head_dim"] == (1, 2, 4, 4)
        assert g["cross_attention_dim"] == 128
        assert g["down_has_attn"] == (True, True, True, False) and g["up_has_attn"] == (False, True, True, True)
        assert infer_vae_channels(synthetic_vae_state_dict((64, 64, 128, 128), lora_rank=rank)) == (64, 64, 128, 128)
    with pytest.raises(KeyError):
        infer_unet_geometry({"foo.weight": torch.zeros(1)})
