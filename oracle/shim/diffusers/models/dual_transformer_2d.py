class DualTransformer2DModel:
    def __init__(self, *a, **k):
        raise NotImplementedError("dual_cross_attention is False for SD-Turbo")
