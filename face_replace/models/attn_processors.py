"""Reference import path `face_replace.models.attn_processors` (reference file of the same name, :1-331): the
processors, `adain` and the two registration helpers, implemented on the B200 kernels."""
from instantrestore_b200.attn_processors import (  # noqa: F401
    ADAIN_EPS,
    AttnProcessor,
    FaceIDAttnProcessor,
    SharedAttnProcessor,
    adain,
    register_attention_processor,
    register_attention_processor_kv_unet,
)

__all__ = ["adain", "AttnProcessor", "FaceIDAttnProcessor", "SharedAttnProcessor", "register_attention_processor",
           "register_attention_processor_kv_unet"]
