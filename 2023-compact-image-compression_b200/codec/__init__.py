"""Mirror of the reference's `codec` package: `from codec.core import Encoder, Decoder`."""
from .core import Decoder, Encoder, decode_batch, encode_batch  # noqa: F401
