from oracle.diffusers024 import Timesteps, TimestepEmbedding  # noqa: F401


class _Unused:
    def __init__(self, *a, **k):
        raise NotImplementedError("not reachable for the SD-Turbo configuration")


GaussianFourierProjection = ImageHintTimeEmbedding = ImageProjection = ImageTimeEmbedding = _Unused
PositionNet = TextImageProjection = TextImageTimeEmbedding = TextTimeEmbedding = _Unused
