"""compact-b200: B200-native implementation of the CompaCT lossless codec hot path.

The directory name is not a Python identifier; import it with
`importlib.import_module("2023-compact-image-compression_b200")` or through the `compact_b200` shim at the
repository root.  Layout:
  csrc/      hand-written sm_100a CUDA kernels + the C ABI (include/compact_b200.h) -> libcompact_b200.so
  codec/     host-side mirror of the reference's `codec.core` (Encoder / Decoder) + batch entry points
  synth.py   seeded synthetic CT / radiograph generators used by tests and bench
  evaluate.py  the reference's compression-ratio study (scripts/evaluate.py) on arrays, CCT column from the GPU
  shard.py   contiguous shards of a batch over the ranks of one box
"""
from . import _native  # noqa: F401
from .codec.core import (DEFAULT_CONFIG, Decoder, Encoder, decode_batch, default_config, encode_batch,  # noqa: F401
                         read_header)

__all__ = ["Encoder", "Decoder", "encode_batch", "decode_batch", "read_header", "default_config", "DEFAULT_CONFIG"]
