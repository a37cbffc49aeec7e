"""Reference import path `face_replace.models.pix2pix_turbo`: the inference surface of `Pix2Pix_Turbo` (reference
:281-343 `forward(c_t, face_embeds=None, conditioning_images=None, valid_indices=None, mask=None,
return_self_attention_maps=False) -> (x_pred, x_conds, attn_maps)`, :242-279 `get_conditioning_keys_values`) is
instantrestore_b200.pipeline.RestorePipeline, built from a checkpoint's state_dict instead of `from_pretrained`."""
from instantrestore_b200.pipeline import ModelFlags, RefCache, RestorePipeline  # noqa: F401

Pix2Pix_Turbo = RestorePipeline

__all__ = ["Pix2Pix_Turbo", "RestorePipeline", "ModelFlags", "RefCache"]
