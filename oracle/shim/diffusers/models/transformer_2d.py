from oracle.diffusers024 import Transformer2DModel  # noqa: F401
