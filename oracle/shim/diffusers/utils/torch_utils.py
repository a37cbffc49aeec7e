from oracle.diffusers024 import fourier_filter  # noqa: F401
