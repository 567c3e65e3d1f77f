// lz_blocks.cuh -- device helpers shared by the expansion kernels: warp-wide copy, LZVN interpreter.
#pragma once
#include "common.cuh"

namespace lzb {

__device__ __forceinline__ void warp_copy(uint8_t *dst, const uint8_t *src, uint64_t n, uint32_t lane) {
    // head: align dst to 16
    uint64_t head = (16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15;
    if (head > n) head = n;
    if (lane < head) dst[lane] = src[lane];
    dst += head; src += head; n -= head;
    uint64_t nv = n / 16;
    if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        const uint4 *s4 = reinterpret_cast<const uint4 *>(src);
        uint4 *d4 = reinterpret_cast<uint4 *>(dst);
        for (uint64_t i = lane; i < nv; i += 32) d4[i] = __ldg(s4 + i);
    } else {
        uint4 *d4 = reinterpret_cast<uint4 *>(dst);
        for (uint64_t i = lane; i < nv; i += 32) {
            const uint8_t *p = src + i * 16;
            uint4 v;
            v.x = ld_u32(p); v.y = ld_u32(p + 4); v.z = ld_u32(p + 8); v.w = ld_u32(p + 12);
            d4[i] = v;
        }
    }
    uint64_t done = nv * 16;
    for (uint64_t i = done + lane; i < n; i += 32) dst[i] = src[i];
}

// LZVN opcode classes (vn/constants.rs:24-72) in closed form.
enum { OP_SML_L, OP_LRG_L, OP_SML_M, OP_LRG_M, OP_PRE_D, OP_SML_D, OP_MED_D, OP_LRG_D, OP_EOS, OP_UDEF, OP_NOP };
__device__ __forceinline__ int vn_op(uint32_t b) {
    uint32_t hi = b >> 4, lo3 = b & 7;
    if (hi == 0xE) return b == 0xE0 ? OP_LRG_L : OP_SML_L;
    if (hi == 0xF) return b == 0xF0 ? OP_LRG_M : OP_SML_M;
    if (hi == 0x7 || hi == 0xD) return OP_UDEF;
    if (hi == 0xA || hi == 0xB) return OP_MED_D;
    if (lo3 == 7) return OP_LRG_D;
    if (lo3 == 6) {
        if (b == 0x06) return OP_EOS;
        if (b == 0x0E || b == 0x16) return OP_NOP;
        if (b < 0x40) return OP_UDEF;
        return OP_PRE_D;
    }
    return OP_SML_D;
}

// LZVN block, interpreted by one lane (vn/vn_core.rs:51-286).  `stream_out` = first output byte of the
// stream, `out_pos` = bytes of the stream already produced, `cap_end` = the stream's dst capacity: a write
// past it is the C-ABI's BufferOverflow (the reference's Vec would grow).  A block that produces more
// than its header announces can only overwrite later output of its own stream, which then fails.
__device__ inline int vn_decode_block(const uint8_t *src, uint64_t src_rest /* bytes from block start to stream end */,
                               uint8_t *stream_out, uint64_t out_pos, uint64_t cap_end) {
    uint32_t n_raw = ld_u32(src + 4), n_payload = ld_u32(src + 8), match_distance = 0;
    uint64_t p = kVnHeaderSize;  // offset inside src
    for (;;) {
        const uint64_t src_len = src_rest - p;
        const uint64_t vlen = src_len < kVnPayloadLimit ? src_len : kVnPayloadLimit;
        const uint64_t out0 = out_pos;
        uint64_t used = 0;
        int res = LZFSE_B200_OK;
        bool eos = false;
        if (vlen < 8) res = LZFSE_B200_PAYLOAD_UNDERFLOW;
        while (res == LZFSE_B200_OK && !eos) {
            const uint8_t *s = src + p + used;
            const uint64_t rem = vlen - used;
            const uint32_t opu = ld_u32(s);
            uint32_t L = 0, M = 0, D = 0, oplen = 0;
            const int op = vn_op(opu & 0xFF);
            switch (op) {
            case OP_SML_L: L = opu & 0xF; oplen = 1; break;
            case OP_LRG_L: L = ((opu >> 8) & 0xFF) + 16; oplen = 2; break;
            case OP_SML_M: M = opu & 0xF; oplen = 1; break;
            case OP_LRG_M: M = ((opu >> 8) & 0xFF) + 16; oplen = 2; break;
            case OP_PRE_D: M = ((opu >> 3) & 7) + 3; L = (opu >> 6) & 3; oplen = 1; break;
            case OP_SML_D: D = ((opu & 7) << 8) | ((opu >> 8) & 0xFF); M = ((opu >> 3) & 7) + 3; L = (opu >> 6) & 3; oplen = 2; break;
            case OP_MED_D: M = (((opu & 7) << 2) | ((opu >> 8) & 3)) + 3; L = (opu >> 3) & 3; D = (opu >> 10) & 0x3FFF; oplen = 3; break;
            case OP_LRG_D: M = ((opu >> 3) & 7) + 3; L = (opu >> 6) & 3; D = (opu >> 8) & 0xFFFF; oplen = 3; break;
            case OP_NOP: oplen = 1; break;
            case OP_EOS:
                if (ld_u64(s) != 0x06ull) res = LZFSE_B200_VN_BAD_PAYLOAD;
                else { used += 8; eos = true; }
                continue;
            default: res = LZFSE_B200_VN_BAD_OPCODE; continue;
            }
            if (rem - oplen < (uint64_t)L + 8) { res = LZFSE_B200_PAYLOAD_UNDERFLOW; continue; }
            if (op == OP_SML_D || op == OP_MED_D || op == OP_LRG_D) match_distance = D;
            if (L) {
                if (out_pos + L > cap_end) { res = LZFSE_B200_BUFFER_OVERFLOW; continue; }
                for (uint32_t t = 0; t < L; t++) stream_out[out_pos + t] = s[oplen + t];
                out_pos += L;
            }
            if (M) {
                if (match_distance == 0 || match_distance > out_pos) { res = LZFSE_B200_BAD_D_VALUE; continue; }
                if (out_pos + M > cap_end) { res = LZFSE_B200_BUFFER_OVERFLOW; continue; }
                uint8_t *q = stream_out + out_pos;
                for (uint32_t t = 0; t < M; t++) q[t] = q[(int64_t)t - (int64_t)match_distance];
                out_pos += M;
            }
            used += oplen + L;
        }
        const uint64_t produced = out_pos - out0;
        if (used > n_payload) return LZFSE_B200_PAYLOAD_UNDERFLOW;
        if (produced > n_raw) return LZFSE_B200_VN_BAD_PAYLOAD;
        n_payload -= (uint32_t)used;
        n_raw -= (uint32_t)produced;
        const bool cycle = src_len > kVnPayloadLimit;
        p += used;
        if (res == LZFSE_B200_OK) {
            if (n_payload != 0) return LZFSE_B200_PAYLOAD_OVERFLOW;
            if (n_raw != 0) return LZFSE_B200_VN_BAD_PAYLOAD;
            return LZFSE_B200_OK;
        }
        if (res == LZFSE_B200_PAYLOAD_UNDERFLOW && cycle) continue;
        return res;
    }
}

// ------------------------------------------------------------------------------------------------
// Streams that consist of ONE small LZVN block -- what the reference's encoder emits for every input of 21..4096
// bytes (encode/frontend_bytes.rs:63-77) -- are expanded by k_expand_vn (decode.cu), a warp per stream with the
// output assembled in shared memory; the in-order expansion kernels skip them.  Everything else (LZVN blocks inside
// longer frames, payloads beyond the interpreter's 0x2000-byte window, too small a destination) takes the one-lane
// interpreter above, which is also the fast kernel's fallback whenever a stream is not well-formed.
// ------------------------------------------------------------------------------------------------
constexpr uint32_t kVnFastRaw = 4096;
__device__ __forceinline__ bool vn_fast_eligible(uint64_t n_stream_blocks, const BlockDesc &bd, uint64_t src_rest, uint64_t cap) {
    return n_stream_blocks == 1 && bd.type == BT_VXN && bd.pad == 0 && bd.n_raw <= kVnFastRaw && src_rest >= kVnHeaderSize + 8 &&
           src_rest - kVnHeaderSize <= kVnPayloadLimit && cap >= bd.n_raw;
}

// Unaligned little-endian loads through aligned words (the bytes up to the next word boundary may lie past `p + n`;
// callers guarantee that word still belongs to the buffer).
__device__ __forceinline__ uint32_t ldg4u(const uint8_t *p) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t r = (uint32_t)a & 3u;
    const uint32_t *q = reinterpret_cast<const uint32_t *>(a - r);
    const uint32_t lo = __ldg(q), hi = r ? __ldg(q + 1) : 0u;
    return __funnelshift_r(lo, hi, r * 8);
}

}  // namespace lzb
