"""`face_replace` — the reference's package path, served by instantrestore_b200.

The reference (snap-research/InstantRestore) is used as `from face_replace.inference.test import Predictor`,
`from face_replace.models.attn_processors import register_attention_processor, SharedAttnProcessor`, with checkpoints
and `config_files/*.yaml` in its own layout. This package keeps those import paths, class names, constructor
arguments and attributes and routes them to the B200 engine (instantrestore_b200); nothing under it computes on the
CPU. Out-of-scope parts of the reference package (training, data, losses) are not provided.
"""
