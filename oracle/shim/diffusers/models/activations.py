from oracle.diffusers024 import get_activation  # noqa: F401
