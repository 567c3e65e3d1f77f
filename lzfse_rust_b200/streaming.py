"""Streaming front doors over the GPU engine: LzfseRingEncoder / LzfseRingDecoder and the reader / writer adaptors.

Mirrors lzfse_rust's streaming API surface -- `LzfseRingEncoder::encode(&mut src, &mut dst) -> (n_raw, n_payload)`
(src/encode/ring_encoder.rs:55-98), `LzfseRingDecoder::decode(&mut src, &mut dst)` (src/decode/ring_decoder.rs:57-90),
`reader()` / `writer()` (src/decode/reader_core.rs:170-187, src/encode/writer.rs) -- on top of the memory-buffer calls.
The reference streams through a 512 KiB ring so that it never holds a whole stream; an LZFSE frame has no stream-level
state beyond its blocks, and the frames it writes are the ones `encode_bytes` writes.  Here a stream is one unit of
GPU work, so these adaptors collect the input (file objects, iterables of chunks) and hand it to the batched engine in
one call: same bytes in and out, same errors, memory proportional to the stream.  No CPU codec is involved."""
import io

from .codec import LzfseDecoder, LzfseEncoder, LzfseError


def _read_all(src):
    if isinstance(src, (bytes, bytearray, memoryview)):
        return bytes(src)
    if hasattr(src, "read"):
        chunks = []
        while True:
            c = src.read(1 << 20)
            if not c:
                break
            chunks.append(c)
        return b"".join(chunks)
    return b"".join(bytes(c) for c in src)


class LzfseRingEncoder:
    """lzfse_rust::LzfseRingEncoder: encode(src, dst) reads `src` to its end and writes one frame to `dst`."""

    def __init__(self, device=0):
        self._enc = LzfseEncoder(device)

    def encode(self, src, dst):
        """Returns (n_raw_bytes, n_payload_bytes) like the reference."""
        raw = _read_all(src)
        out = bytearray()
        self._enc.encode_bytes(raw, out)
        dst.write(out)
        return len(raw), len(out)

    def encode_bytes(self, src, dst):
        return self._enc.encode_bytes(src, dst)

    def writer(self, inner):
        return LzfseWriter(self, inner)

    def close(self):
        self._enc.close()


class LzfseRingDecoder:
    """lzfse_rust::LzfseRingDecoder: decode(src, dst) reads one frame from `src` to its end and writes the bytes to `dst`."""

    def __init__(self, device=0):
        self._dec = LzfseDecoder(device)

    def decode(self, src, dst):
        """Returns (n_raw_bytes, n_payload_bytes) like the reference; raises LzfseError with the reference's error."""
        frame = _read_all(src)
        out = bytearray()
        self._dec.decode_bytes(frame, out)
        dst.write(out)
        return len(out), len(frame)

    def decode_bytes(self, src, dst):
        return self._dec.decode_bytes(src, dst)

    def reader(self, inner):
        return LzfseReader(self, inner)

    def close(self):
        self._dec.close()


class LzfseWriter(io.RawIOBase):
    """`Write` adaptor (lzfse_rust::LzfseWriter): bytes written are encoded into one frame when finalize() is called;
    like the reference, dropping the writer without finalize() loses the frame's tail."""

    def __init__(self, encoder, inner):
        super().__init__()
        self._encoder, self._inner, self._buf, self._done = encoder, inner, bytearray(), False

    def writable(self):
        return True

    def write(self, b):
        if self._done:
            raise ValueError("write after finalize")
        self._buf += bytes(b)
        return len(b)

    def finalize(self):
        """Encodes what was written, writes the frame to the inner writer and returns it (the reference's finalize)."""
        if not self._done:
            self._done = True
            out = bytearray()
            self._encoder.encode_bytes(bytes(self._buf), out)
            self._inner.write(out)
            self._buf = bytearray()
        return self._inner


class LzfseReader(io.RawIOBase):
    """`Read` adaptor (lzfse_rust::LzfseReader): the inner reader's frame is decoded on first use; a corrupt frame raises
    LzfseError at the first read, where the reference's reader reports it when it reaches the bad block."""

    def __init__(self, decoder, inner):
        super().__init__()
        self._decoder, self._inner, self._out, self._pos = decoder, inner, None, 0

    def readable(self):
        return True

    def _fill(self):
        if self._out is None:
            out = bytearray()
            self._decoder.decode_bytes(_read_all(self._inner), out)
            self._out = bytes(out)

    def readinto(self, b):
        self._fill()
        n = min(len(b), len(self._out) - self._pos)
        b[:n] = self._out[self._pos:self._pos + n]
        self._pos += n
        return n

    def into_inner(self):
        return self._inner


__all__ = ["LzfseRingEncoder", "LzfseRingDecoder", "LzfseWriter", "LzfseReader", "LzfseError"]
