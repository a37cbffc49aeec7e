class AdaGroupNorm:
    def __init__(self, *a, **k):
        raise NotImplementedError("resnet_time_scale_shift is 'default' for SD-Turbo")
