//! Raw bindings to `include/lzfse_b200.h`.  Pointers and sizes only; see the header for the contracts.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
pub struct lzfse_b200_decoder { _p: [u8; 0] }
#[repr(C)]
pub struct lzfse_b200_encoder { _p: [u8; 0] }

// enum lzfse_b200_status
pub const LZFSE_B200_OK: c_int = 0;
pub const LZFSE_B200_BAD_BLOCK: c_int = 1;
pub const LZFSE_B200_BAD_BITSTREAM: c_int = 2;
pub const LZFSE_B200_BAD_D_VALUE: c_int = 3;
pub const LZFSE_B200_BAD_READER_STATE: c_int = 4;
pub const LZFSE_B200_BUFFER_OVERFLOW: c_int = 5;
pub const LZFSE_B200_PAYLOAD_OVERFLOW: c_int = 6;
pub const LZFSE_B200_PAYLOAD_UNDERFLOW: c_int = 7;
pub const LZFSE_B200_FSE_BASE: c_int = 16; // + FseErrorKind discriminant (src/fse/error_kind.rs:9-39)
pub const LZFSE_B200_VN_BASE: c_int = 32;  // + VnErrorKind discriminant (src/vn/error_kind.rs:9-16)
pub const LZFSE_B200_INVALID_ARGUMENT: c_int = 64;
pub const LZFSE_B200_NO_DEVICE: c_int = 65;
pub const LZFSE_B200_CUDA_ERROR: c_int = 66;
pub const LZFSE_B200_OUT_OF_MEMORY: c_int = 67;

extern "C" {
    pub fn lzfse_b200_version() -> *const c_char;
    pub fn lzfse_b200_status_string(status: c_int) -> *const c_char;
    pub fn lzfse_b200_decoder_last_error(d: *const lzfse_b200_decoder) -> *const c_char;
    pub fn lzfse_b200_encoder_last_error(e: *const lzfse_b200_encoder) -> *const c_char;

    pub fn lzfse_b200_decoder_create(cuda_device: c_int, out: *mut *mut lzfse_b200_decoder) -> c_int;
    pub fn lzfse_b200_decoder_destroy(d: *mut lzfse_b200_decoder);
    pub fn lzfse_b200_decode_bytes(d: *mut lzfse_b200_decoder, src: *const u8, src_len: usize, dst: *mut u8, dst_cap: usize,
                                   dst_len: *mut usize) -> c_int;
    pub fn lzfse_b200_decode_batch_device(d: *mut lzfse_b200_decoder, src_base: *const u8, src_off: *const u64, src_len: *const u64,
                                          dst_base: *mut u8, dst_off: *const u64, dst_cap: *const u64, out_len: *mut u64,
                                          status: *mut i32, n: usize, cuda_stream: *mut c_void) -> c_int;
    pub fn lzfse_b200_decode_batch_host(d: *mut lzfse_b200_decoder, src_base: *const u8, src_off: *const u64, src_len: *const u64,
                                        dst_base: *mut u8, dst_off: *const u64, dst_cap: *const u64, out_len: *mut u64,
                                        status: *mut i32, n: usize) -> c_int;
    pub fn lzfse_b200_decode_batch_device_async(d: *mut lzfse_b200_decoder, src_base: *const u8, src_off: *const u64,
                                                src_len: *const u64, dst_base: *mut u8, dst_off: *const u64, dst_cap: *const u64,
                                                out_len: *mut u64, status: *mut i32, n: usize, cuda_stream: *mut c_void) -> c_int;
    pub fn lzfse_b200_decoder_sync(d: *mut lzfse_b200_decoder) -> c_int;
    pub fn lzfse_b200_decode_probe_batch_device(d: *mut lzfse_b200_decoder, src_base: *const u8, src_off: *const u64,
                                                src_len: *const u64, raw_len: *mut u64, n_blocks: *mut u32, status: *mut i32,
                                                n: usize, cuda_stream: *mut c_void) -> c_int;
    pub fn lzfse_b200_decode_probe_batch_host(d: *mut lzfse_b200_decoder, src_base: *const u8, src_off: *const u64,
                                              src_len: *const u64, raw_len: *mut u64, n_blocks: *mut u32, status: *mut i32,
                                              n: usize) -> c_int;
    pub fn lzfse_b200_decode_prefix_batch_device(d: *mut lzfse_b200_decoder, src_base: *const u8, src_off: *const u64,
                                                 src_len: *const u64, dst_base: *mut u8, dst_off: *const u64, limit: *const u64,
                                                 out_len: *mut u64, status: *mut i32, more: *mut u8, n: usize,
                                                 cuda_stream: *mut c_void) -> c_int;
    pub fn lzfse_b200_decode_prefix_batch_host(d: *mut lzfse_b200_decoder, src_base: *const u8, src_off: *const u64,
                                               src_len: *const u64, dst_base: *mut u8, dst_off: *const u64, limit: *const u64,
                                               out_len: *mut u64, status: *mut i32, more: *mut u8, n: usize) -> c_int;
    pub fn lzfse_b200_decoder_last_launches(d: *const lzfse_b200_decoder) -> u64;
    pub fn lzfse_b200_decoder_set_timing(d: *mut lzfse_b200_decoder, enabled: c_int);
    pub fn lzfse_b200_decoder_last_stage_ms(d: *const lzfse_b200_decoder, stage_ms: *mut f32, cap: c_int) -> c_int;

    pub fn lzfse_b200_encoder_create(cuda_device: c_int, out: *mut *mut lzfse_b200_encoder) -> c_int;
    pub fn lzfse_b200_encoder_destroy(e: *mut lzfse_b200_encoder);
    pub fn lzfse_b200_encode_bound(src_len: usize) -> usize;
    pub fn lzfse_b200_encode_bound_strict(src_len: usize) -> usize;
    pub fn lzfse_b200_encode_bytes(e: *mut lzfse_b200_encoder, src: *const u8, src_len: usize, dst: *mut u8, dst_cap: usize,
                                   dst_len: *mut usize) -> c_int;
    pub fn lzfse_b200_encode_batch_device(e: *mut lzfse_b200_encoder, src_base: *const u8, src_off: *const u64, src_len: *const u64,
                                          dst_base: *mut u8, dst_off: *const u64, dst_cap: *const u64, out_len: *mut u64,
                                          status: *mut i32, n: usize, cuda_stream: *mut c_void) -> c_int;
    pub fn lzfse_b200_encode_batch_host(e: *mut lzfse_b200_encoder, src_base: *const u8, src_off: *const u64, src_len: *const u64,
                                        dst_base: *mut u8, dst_off: *const u64, dst_cap: *const u64, out_len: *mut u64,
                                        status: *mut i32, n: usize) -> c_int;
    pub fn lzfse_b200_encode_batch_device_async(e: *mut lzfse_b200_encoder, src_base: *const u8, src_off: *const u64,
                                                src_len: *const u64, dst_base: *mut u8, dst_off: *const u64, dst_cap: *const u64,
                                                out_len: *mut u64, status: *mut i32, n: usize, cuda_stream: *mut c_void) -> c_int;
    pub fn lzfse_b200_encoder_sync(e: *mut lzfse_b200_encoder) -> c_int;
    pub fn lzfse_b200_encoder_last_launches(e: *const lzfse_b200_encoder) -> u64;
    pub fn lzfse_b200_encoder_set_timing(e: *mut lzfse_b200_encoder, enabled: c_int);
    pub fn lzfse_b200_encoder_last_stage_ms(e: *const lzfse_b200_encoder, stage_ms: *mut f32, cap: c_int) -> c_int;
}
