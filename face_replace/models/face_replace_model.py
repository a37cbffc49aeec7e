"""Reference import path `face_replace.models.face_replace_model` (reference :8-45): `FaceReplaceModel(cfg, full_cfg)`
whose `.net` runs the single-step restoration forward — here on the B200 engine."""
from instantrestore_b200.inference import FaceReplaceModel  # noqa: F401

__all__ = ["FaceReplaceModel"]
