"""lzfse_rust_b200 -- B200-native batched LZFSE codec behind lzfse_rust's encode_bytes/decode_bytes API.

The codec itself is hand-written CUDA for sm_100a (csrc/) reached through the C-ABI in
include/lzfse_b200.h; this package is the thin host-side mirror of the reference's Rust API.
There is no CPU implementation in here: creating a decoder/encoder without a CUDA device raises.
"""
from .codec import LzfseDecoder, LzfseEncoder, LzfseError, STATUS_NAMES, decode_bytes, encode_bytes
from .sharding import shard_ranges
from .streaming import LzfseReader, LzfseRingDecoder, LzfseRingEncoder, LzfseWriter

__all__ = ["LzfseDecoder", "LzfseEncoder", "LzfseError", "STATUS_NAMES", "decode_bytes", "encode_bytes", "shard_ranges",
           "LzfseRingDecoder", "LzfseRingEncoder", "LzfseReader", "LzfseWriter"]
__version__ = "0.2.0"
