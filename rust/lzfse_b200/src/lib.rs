//! Safe wrapper over `lzfse_b200_sys` with the names and contracts of lzfse_rust v0.2.0's memory-buffer engine:
//! `LzfseEncoder::encode_bytes` (src/encode/encoder.rs:49), `LzfseDecoder::decode_bytes` (src/decode/decoder.rs:61),
//! the free functions (src/encode/mod.rs:58, src/decode/mod.rs:49) and `Error` (src/error/mod.rs:40-61).
//! New: `encode_batch` / `decode_batch` over many independent streams in one GPU call.
//! There is no CPU fallback: without a CUDA device `default()` panics and `try_new` returns `Error::Io`.
use lzfse_b200_sys as sys;
use std::fmt;
use std::io;

/// lzfse_rust::FseErrorKind (src/fse/error_kind.rs:9-39), same discriminants.
#[derive(Copy, Clone, Debug, PartialEq, Eq)]
#[repr(u8)]
pub enum FseErrorKind {
    BadLiteralBits, BadLiteralCount, BadLiteralPayload, BadLiteralState, BadLmdBits, BadLmdCount, BadLmdPayload, BadLmdState,
    BadPayloadCount, BadRawByteCount, BadReaderState, BadWeightPayload, BadWeightPayloadCount, WeightPayloadOverflow,
    WeightPayloadUnderflow,
}
/// lzfse_rust::VnErrorKind (src/vn/error_kind.rs:9-16).
#[derive(Copy, Clone, Debug, PartialEq, Eq)]
#[repr(u8)]
pub enum VnErrorKind { BadPayloadCount, BadPayload, BadOpcode }

/// lzfse_rust::Error (src/error/mod.rs:40-61).
#[derive(Debug)]
pub enum Error {
    Io(io::Error),
    BufferOverflow,
    BadBlock(u32),
    BadBitStream,
    BadDValue,
    BadReaderState,
    Fse(FseErrorKind),
    Vn(VnErrorKind),
    PayloadOverflow,
    PayloadUnderflow,
}
pub type Result<T> = std::result::Result<T, Error>;

const FSE_KINDS: [FseErrorKind; 15] = [
    FseErrorKind::BadLiteralBits, FseErrorKind::BadLiteralCount, FseErrorKind::BadLiteralPayload, FseErrorKind::BadLiteralState,
    FseErrorKind::BadLmdBits, FseErrorKind::BadLmdCount, FseErrorKind::BadLmdPayload, FseErrorKind::BadLmdState,
    FseErrorKind::BadPayloadCount, FseErrorKind::BadRawByteCount, FseErrorKind::BadReaderState, FseErrorKind::BadWeightPayload,
    FseErrorKind::BadWeightPayloadCount, FseErrorKind::WeightPayloadOverflow, FseErrorKind::WeightPayloadUnderflow,
];

impl Error {
    /// Inverse of `enum lzfse_b200_status` (include/lzfse_b200.h).
    pub fn from_status(st: i32) -> Error {
        match st {
            1 => Error::BadBlock(0),
            2 => Error::BadBitStream,
            3 => Error::BadDValue,
            4 => Error::BadReaderState,
            5 => Error::BufferOverflow,
            6 => Error::PayloadOverflow,
            7 => Error::PayloadUnderflow,
            16..=30 => Error::Fse(FSE_KINDS[(st - 16) as usize]),
            32 => Error::Vn(VnErrorKind::BadPayloadCount),
            33 => Error::Vn(VnErrorKind::BadPayload),
            34 => Error::Vn(VnErrorKind::BadOpcode),
            _ => Error::Io(io::Error::new(io::ErrorKind::Other, format!("lzfse_b200 status {}", st))),
        }
    }
}
impl fmt::Display for Error {
    fn fmt(&self, f: &mut fmt::Formatter<'_>) -> fmt::Result { write!(f, "{:?}", self) }
}
impl std::error::Error for Error {}
impl From<Error> for io::Error {
    fn from(e: Error) -> io::Error {
        match e { Error::Io(e) => e, e => io::Error::new(io::ErrorKind::InvalidData, e) }
    }
}

/// One stream of a batch: its bytes on success, the reference's error otherwise.
pub type StreamResult = Result<Vec<u8>>;

fn offsets(lens: &[u64]) -> Vec<u64> {
    let mut off = Vec::with_capacity(lens.len());
    let mut acc = 0u64;
    for l in lens { off.push(acc); acc += *l; }
    off
}

pub struct LzfseDecoder(*mut sys::lzfse_b200_decoder);
unsafe impl Send for LzfseDecoder {}

impl LzfseDecoder {
    pub fn try_new(cuda_device: i32) -> Result<Self> {
        let mut h = std::ptr::null_mut();
        let rc = unsafe { sys::lzfse_b200_decoder_create(cuda_device, &mut h) };
        if rc != 0 { return Err(Error::from_status(rc)); }
        Ok(Self(h))
    }

    /// Same contract as lzfse_rust::LzfseDecoder::decode_bytes: appends to `dst`, returns the bytes appended.
    /// (A match may not reach into bytes that were in `dst` before the call; the reference allows it, lz/writer.rs:156-157.)
    pub fn decode_bytes(&mut self, src: &[u8], dst: &mut Vec<u8>) -> Result<u64> {
        let (mut raw, mut nb, mut st) = (0u64, 0u32, 0i32);
        let (off, len) = (0u64, src.len() as u64);
        let rc = unsafe { sys::lzfse_b200_decode_probe_batch_host(self.0, src.as_ptr(), &off, &len, &mut raw, &mut nb, &mut st, 1) };
        if rc != 0 { return Err(Error::from_status(rc)); }
        if st != 0 { return Err(Error::from_status(st)); }
        let cap = raw as usize;
        dst.reserve(cap);
        let mut n = 0usize;
        let rc = unsafe { sys::lzfse_b200_decode_bytes(self.0, src.as_ptr(), src.len(), dst.as_mut_ptr().add(dst.len()), cap, &mut n) };
        if rc != 0 { return Err(Error::from_status(rc)); }
        unsafe { dst.set_len(dst.len() + n) };
        Ok(n as u64)
    }

    /// n independent frames in one GPU call.  A failing frame does not disturb the others.
    pub fn decode_batch(&mut self, frames: &[&[u8]]) -> Result<Vec<StreamResult>> {
        let n = frames.len();
        let src_len: Vec<u64> = frames.iter().map(|f| f.len() as u64).collect();
        let src_off = offsets(&src_len);
        let src: Vec<u8> = frames.concat();
        let (mut raw, mut nb, mut st) = (vec![0u64; n], vec![0u32; n], vec![0i32; n]);
        let rc = unsafe { sys::lzfse_b200_decode_probe_batch_host(self.0, src.as_ptr(), src_off.as_ptr(), src_len.as_ptr(),
                                                                  raw.as_mut_ptr(), nb.as_mut_ptr(), st.as_mut_ptr(), n) };
        if rc != 0 { return Err(Error::from_status(rc)); }
        // what a frame announces is untrusted: cap it at what a frame of its size can produce
        let cap: Vec<u64> = raw.iter().zip(&src_len).map(|(r, l)| (*r).min(l * 65536 + 65536)).collect();
        let dst_off = offsets(&cap);
        let mut dst = vec![0u8; cap.iter().sum::<u64>() as usize];
        let mut out_len = vec![0u64; n];
        let rc = unsafe { sys::lzfse_b200_decode_batch_host(self.0, src.as_ptr(), src_off.as_ptr(), src_len.as_ptr(), dst.as_mut_ptr(),
                                                            dst_off.as_ptr(), cap.as_ptr(), out_len.as_mut_ptr(), st.as_mut_ptr(), n) };
        if rc != 0 { return Err(Error::from_status(rc)); }
        Ok((0..n).map(|i| if st[i] == 0 { Ok(dst[dst_off[i] as usize..(dst_off[i] + out_len[i]) as usize].to_vec()) }
                          else { Err(Error::from_status(st[i])) }).collect())
    }
}
impl Default for LzfseDecoder {
    fn default() -> Self { Self::try_new(0).expect("no usable CUDA device: the B200 path has no CPU fallback") }
}
impl Drop for LzfseDecoder {
    fn drop(&mut self) { unsafe { sys::lzfse_b200_decoder_destroy(self.0) } }
}
impl fmt::Debug for LzfseDecoder {
    fn fmt(&self, f: &mut fmt::Formatter<'_>) -> fmt::Result { f.debug_struct("LzfseDecoder").finish() }
}

pub struct LzfseEncoder(*mut sys::lzfse_b200_encoder);
unsafe impl Send for LzfseEncoder {}

impl LzfseEncoder {
    pub fn try_new(cuda_device: i32) -> Result<Self> {
        let mut h = std::ptr::null_mut();
        let rc = unsafe { sys::lzfse_b200_encoder_create(cuda_device, &mut h) };
        if rc != 0 { return Err(Error::from_status(rc)); }
        Ok(Self(h))
    }

    /// Same contract as lzfse_rust::LzfseEncoder::encode_bytes: appends one frame to `dst`, returns the bytes appended.
    pub fn encode_bytes(&mut self, src: &[u8], dst: &mut Vec<u8>) -> io::Result<u64> {
        let cap = unsafe { sys::lzfse_b200_encode_bound(src.len()) };
        dst.reserve(cap);
        let mut n = 0usize;
        let rc = unsafe { sys::lzfse_b200_encode_bytes(self.0, src.as_ptr(), src.len(), dst.as_mut_ptr().add(dst.len()), cap, &mut n) };
        if rc != 0 { return Err(Error::from_status(rc).into()); }
        unsafe { dst.set_len(dst.len() + n) };
        Ok(n as u64)
    }

    /// n independent inputs -> n independent frames, each byte-identical to what `encode_bytes` gives for it.
    pub fn encode_batch(&mut self, inputs: &[&[u8]]) -> io::Result<Vec<Vec<u8>>> {
        let n = inputs.len();
        let src_len: Vec<u64> = inputs.iter().map(|f| f.len() as u64).collect();
        let src_off = offsets(&src_len);
        let src: Vec<u8> = inputs.concat();
        let cap: Vec<u64> = src_len.iter().map(|l| unsafe { sys::lzfse_b200_encode_bound(*l as usize) } as u64).collect();
        let dst_off = offsets(&cap);
        let mut dst = vec![0u8; cap.iter().sum::<u64>() as usize];
        let (mut out_len, mut st) = (vec![0u64; n], vec![0i32; n]);
        let rc = unsafe { sys::lzfse_b200_encode_batch_host(self.0, src.as_ptr(), src_off.as_ptr(), src_len.as_ptr(), dst.as_mut_ptr(),
                                                            dst_off.as_ptr(), cap.as_ptr(), out_len.as_mut_ptr(), st.as_mut_ptr(), n) };
        if rc != 0 { return Err(Error::from_status(rc).into()); }
        if let Some(bad) = st.iter().find(|s| **s != 0) { return Err(Error::from_status(*bad).into()); }
        Ok((0..n).map(|i| dst[dst_off[i] as usize..(dst_off[i] + out_len[i]) as usize].to_vec()).collect())
    }
}
impl Default for LzfseEncoder {
    fn default() -> Self { Self::try_new(0).expect("no usable CUDA device: the B200 path has no CPU fallback") }
}
impl Drop for LzfseEncoder {
    fn drop(&mut self) { unsafe { sys::lzfse_b200_encoder_destroy(self.0) } }
}
impl fmt::Debug for LzfseEncoder {
    fn fmt(&self, f: &mut fmt::Formatter<'_>) -> fmt::Result { f.debug_struct("LzfseEncoder").finish() }
}

/// lzfse_rust::decode_bytes (src/decode/mod.rs:49).
pub fn decode_bytes(src: &[u8], dst: &mut Vec<u8>) -> Result<u64> { LzfseDecoder::default().decode_bytes(src, dst) }
/// lzfse_rust::encode_bytes (src/encode/mod.rs:58).
pub fn encode_bytes(src: &[u8], dst: &mut Vec<u8>) -> io::Result<u64> { LzfseEncoder::default().encode_bytes(src, dst) }

#[cfg(test)]
mod tests {
    use super::*;

    // src/encode/mod.rs:50-54 (doc test of the reference)
    #[test]
    fn doc_frame() {
        let mut enc = Vec::new();
        assert_eq!(encode_bytes(b"test", &mut enc).unwrap(), 16);
        assert_eq!(enc, [0x62, 0x76, 0x78, 0x2d, 0x04, 0x00, 0x00, 0x00, 0x74, 0x65, 0x73, 0x74, 0x62, 0x76, 0x78, 0x24]);
        let mut dec = b"keep".to_vec();
        assert_eq!(decode_bytes(&enc, &mut dec).unwrap(), 4);
        assert_eq!(dec, b"keeptest");
    }

    #[test]
    fn batch_round_trip() {
        let a = vec![7u8; 100_000];
        let b: Vec<u8> = (0..70_000u32).map(|i| (i * 2654435761u32 >> 24) as u8).collect();
        let frames = LzfseEncoder::default().encode_batch(&[&a, &b, b""]).unwrap();
        let refs: Vec<&[u8]> = frames.iter().map(|f| &f[..]).collect();
        let outs = LzfseDecoder::default().decode_batch(&refs).unwrap();
        assert_eq!(outs[0].as_ref().unwrap(), &a);
        assert_eq!(outs[1].as_ref().unwrap(), &b);
        assert!(outs[2].as_ref().unwrap().is_empty());
    }
}
