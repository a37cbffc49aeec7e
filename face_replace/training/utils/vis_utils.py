"""Reference import path `face_replace.training.utils.vis_utils`: only `tensor2im` (reference :14-23) is on the
inference path."""
from instantrestore_b200.inference import tensor2im  # noqa: F401

__all__ = ["tensor2im"]
