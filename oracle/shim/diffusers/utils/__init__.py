import logging as _pylogging

import torch

USE_PEFT_BACKEND = True  # peft is installed in the reference environment (environment.yaml:193)


class BaseOutput:
    pass


def deprecate(*args, **kwargs):
    pass


class _Logging:
    @staticmethod
    def get_logger(name):
        return _pylogging.getLogger(name)


logging = _Logging()


def scale_lora_layers(model, weight):
    pass  # lora scale is 1.0 on the reference path (no "scale" in cross_attention_kwargs)


def unscale_lora_layers(model, weight=None):
    pass


def is_torch_version(op, version):
    return True
