"""molly_b200 -- B200-native (sm_100a) implementation of SeedLLM/molly's omics-embedding hot path:
ESM-2 / nucleotide-transformer encoder forward -> encoder-to-LLM projector -> scatter into the Qwen3 ``inputs_embeds``,
behind the reference's own boundary ``OmicsOne.process_omic_sequences`` (reference ``src/model/omics_one.py:49-136``).

Importing the package does not need a GPU; using it does, and it fails loudly (no CPU fallback) when
``libmolly_b200.so`` has not been built (``__graft_entry__.build()``).
"""
from .config import EncoderConfig  # noqa: F401

__all__ = ["EncoderConfig", "FastOmicsPath", "PackedEncoder"]


def __getattr__(name):
    # lazy: these import torch custom-op registration and load the CUDA library
    if name == "FastOmicsPath":
        from .omics_path import FastOmicsPath
        return FastOmicsPath
    if name == "PackedEncoder":
        from .packing import PackedEncoder
        return PackedEncoder
    raise AttributeError(name)
