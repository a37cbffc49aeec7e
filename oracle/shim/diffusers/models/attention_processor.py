from typing import Union

from oracle.diffusers024 import Attention, DefaultAttnProcessor as AttnProcessor  # noqa: F401


class AttnAddedKVProcessor:  # never instantiated for the SD-Turbo configuration
    pass


class AttnAddedKVProcessor2_0:
    pass


ADDED_KV_ATTENTION_PROCESSORS = (AttnAddedKVProcessor, AttnAddedKVProcessor2_0)
CROSS_ATTENTION_PROCESSORS = (AttnProcessor,)
AttentionProcessor = Union[AttnProcessor, AttnAddedKVProcessor]
