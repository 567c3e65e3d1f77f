// encode.cu -- batched LZFSE encode kernels for sm_100a + the encoder half of the C-ABI.
//
// The encoder reproduces lzfse_rust's `encode_bytes` bit for bit (same parse, same blocks, same frames),
// which is possible on a GPU because the only sequential piece of the reference's front end is the lazy
// match arbitration: the hash table is fed every position exactly once and in order
// (encode/frontend_bytes.rs:185-207,336-344), so the candidates a position sees do not depend on the parse.
//
// Pipeline (one CUDA stream):
//   k_enc_prep          thread / stream    block-type policy (frontend_bytes.rs:63-77), scratch sizing, which front end
//   k_exclusive_scan                       per-stream scratch bases (decode.cu)
//   bvx2 streams <= 64 KiB:  k_enc_find    CTA / stream       find_match for every position in shared memory -> one word per position
//   bvx2 streams  > 64 KiB:  k_long_heads / carry / chain / find (encode_long.cuh)   the same words, hash chain over the stream in HBM
//   every bvx2 stream:       k_long_replay / stitch_a / stitch_b (encode_long.cuh)    the sequential front end (backward limit,
//                            Match::select) per 4 Ki-position segment, speculative, stitched where the states meet
//                            k_long_seg_stats / blocks / write_packs                  Buffer::push: packs and block records
//   LZVN-sized inputs:       k_enc_parse   warp / stream      hash table in HBM, 32 positions per step, LZVN opcodes
//   k_enc_fse_blocks    warp / block       histogram, normalize_m1, weight varints, encode tables, 4-state literal stream,
//                                          L/M/D stream, bvx2 header
//   k_enc_assemble (+ k_long_copy)         compaction of the blocks into the caller's frame + bvx$
// (LZB_ENC_SEG=0: k_enc_replay, one thread per <= 64 KiB stream, instead of the segments; LZB_ENC_LONG=0: k_enc_parse for > 64 KiB.)
#include <cstdio>
#include <cstdlib>
#include <new>
#include <string>
#include <type_traits>

#include "common.cuh"
#include "host_util.h"

namespace lzb {

void launch_exclusive_scan(StreamCounts *counts, size_t n, StreamCounts *totals, const uint32_t *work, cudaStream_t s);  // decode.cu

constexpr uint32_t kEmptyIdx = 0xC0C0C0C0u;  // history reset marker (encode/history.rs:72-83: any idx whose distance is out of range)
// A bucket is one 32-byte sector: 4 positions (newest first) followed by the 4 source bytes found at each of them --
// the reference's Item {val, idx} (encode/history.rs:144-154).  Keeping the values next to the positions means a
// candidate is accepted or rejected without touching the source: four random 32-byte source sectors per position
// were most of the parse kernel's HBM traffic.
constexpr uint32_t kBucketWords = 2 * kHashWidth;
constexpr uint32_t kTableWords = (1u << kHashBits) * kBucketWords;
constexpr uint32_t kFastMaxLen = 65536;  // streams up to this long take the shared-memory path (positions fit 16 bits)
#ifndef LZB_LONG_RSEG
#define LZB_LONG_RSEG 4096
#endif
constexpr uint32_t kLongCSeg = 32768, kLongRSeg = LZB_LONG_RSEG;  // longer bvx2 streams: chain pieces / replay segments (encode_long.cuh)

enum StreamKind : uint32_t { SK_RAW = 0, SK_VN = 1, SK_FSE = 2 };

// Per-stream scratch sizing.  A front-end match becomes 1 + M/2359 packs plus L/315 literal-only packs
// (fse/buffer.rs:45-97), and matches are at least 4 bytes apart.
__host__ __device__ inline uint64_t pack_cap(uint64_t len) { return len / 4 + len / 315 + len / 2359 + len / 2048 + 32; }  // + one split pack per block
__host__ __device__ inline uint64_t block_cap(uint64_t len) { return pack_cap(len) / kLmdsPerBlock + len / kLiteralsPerBlock + 3; }
// Hard bound of one bvx2 block: header + weights + 10 bits per (padded) literal + 8 + 54 bits per pack.
__host__ __device__ inline uint64_t block_bound(uint64_t n_lits, uint64_t n_packs) {
    return 32 + 630 + ((n_lits + 3) / 4 * 4 * 10 + 7) / 8 + 8 + (n_packs * 54 + 7) / 8 + 32;
}
__host__ __device__ inline uint64_t out_cap(uint64_t len) {
    return block_cap(len) * 768 + (len + 4 * block_cap(len)) * 10 / 8 + pack_cap(len) * 54 / 8 + 1024;
}

struct EncStream {       // per stream, written by prep / parse, read by assemble
    uint32_t kind;       // StreamKind
    uint32_t n_blocks;   // FSE blocks produced
    uint32_t vn_size;    // LZVN: bytes of the finished block (header included) in the out scratch
    uint32_t fast;       // 1: parsed by k_enc_find + k_enc_replay (bvx2 streams of <= kFastMaxLen bytes), 2: by the k_long_* kernels
                         // (longer bvx2 streams, encode_long.cuh), 0: by k_enc_parse
    uint32_t cseg_base, n_cseg;  // long streams: chain pieces and replay segments (indices into the batch's lists)
    uint32_t rseg_base, n_rseg;
    uint32_t seg, pad;   // seg = 1: the sequential front end runs per segment (k_long_replay + stitch) instead of per stream
    uint64_t long_off;   // long streams: element offset of the stream's prev[] array
};

struct EncBlock {        // one bvx2 block to encode (compact list, any order)
    uint64_t pack_off;   // element offset into the pack scratch
    uint64_t lit_off;    // byte offset into the literal scratch
    uint64_t out_off;    // byte offset into the out scratch
    uint64_t src_pos;    // gather != 0: absolute offset inside src_base of the first byte the block covers
    uint32_t n_packs, n_lits, n_match_bytes;
    uint32_t out_size;   // filled by k_enc_fse_blocks
    uint32_t gather;     // 1: the literal bytes are not in the literal scratch yet (k_enc_fse_blocks collects them from the source)
    uint32_t pad;        // 1: block of a long stream, copied into the frame by k_long_copy (dst_pos set by k_enc_assemble)
    uint64_t dst_pos;    // byte offset of the finished block inside dst_base
};

// ------------------------------------------------------------------------------------------------
// prep: policy + sizing.  counts[i] = {packs, literal bytes, block slots, out bytes}
// ------------------------------------------------------------------------------------------------
__global__ void k_enc_prep(const uint64_t *__restrict__ src_len, size_t n, EncStream *streams, StreamCounts *counts, int32_t *status,
                           uint32_t *path_counts /* [0] fast streams, [1] streams for k_enc_parse, [2] long streams, [3] replay segments,
                                                    [4] chain pieces, [5] streams replayed by segments, [6..7] prev[] elements (u64) */,
                           uint32_t *long_list, uint32_t *seg_list, int allow_fast, int allow_long, int allow_seg) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t len = src_len[i];
    EncStream st;
    st.n_blocks = 0; st.vn_size = 0; st.fast = 0;
    st.cseg_base = st.n_cseg = st.rseg_base = st.n_rseg = 0; st.long_off = 0; st.seg = 0; st.pad = 0;
    StreamCounts c = {0, 0, 0, 0};
    status[i] = LZFSE_B200_OK;
    if (len > 0x7FFFFFFFull) {  // BLOCK_GUIDE repositioning (frontend_bytes.rs:348-375) is out of scope
        st.kind = SK_RAW;
        status[i] = LZFSE_B200_INVALID_ARGUMENT;
    } else if (len > kVnCutoff) {
        st.kind = SK_FSE;
        st.fast = allow_fast && len <= kFastMaxLen;
        if (allow_long && len > kFastMaxLen) {
            st.fast = 2; st.seg = 1;
            st.n_cseg = ((uint32_t)len - 3 + kLongCSeg - 1) / kLongCSeg;
            long_list[atomicAdd(&path_counts[2], 1u)] = (uint32_t)i;
            st.cseg_base = atomicAdd(&path_counts[4], st.n_cseg);
            st.long_off = atomicAdd(reinterpret_cast<unsigned long long *>(path_counts + 6), (unsigned long long)((len + 31) & ~15ull));
        } else if (st.fast == 1 && allow_seg) st.seg = 1;
        if (st.seg) {
            st.n_rseg = ((uint32_t)len - 3 + kLongRSeg - 1) / kLongRSeg;
            st.rseg_base = atomicAdd(&path_counts[3], st.n_rseg);
            seg_list[atomicAdd(&path_counts[5], 1u)] = (uint32_t)i;
        }
        c.n_blocks = pack_cap(len);                  // packs
        c.n_fse = (len + 31) & ~15ull;               // literal bytes
        c.n_literals = block_cap(len);               // block slots
        c.n_lmds = (out_cap(len) + 15) & ~15ull;     // out bytes
    } else if (len > kRawCutoff) {
        st.kind = SK_VN;
        c.n_lmds = (out_cap(len) + 15) & ~15ull;
    } else {
        st.kind = SK_RAW;
    }
    streams[i] = st;
    counts[i] = c;
    if (st.kind != SK_RAW && st.fast != 2) atomicAdd(&path_counts[st.fast ? 0 : 1], 1u);
}
__global__ void k_enc_publish_counts(const uint32_t *path_counts, uint32_t *host_out) { for (int k = 0; k < 8; k++) host_out[k] = path_counts[k]; }

// ------------------------------------------------------------------------------------------------
// parse: warp per stream
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t hash_u(uint32_t val, bool vn) {  // fse/object.rs:38-43, vn/object.rs:33-47
    if (vn) val &= 0x00FFFFFFu;
    return (val * 0x9E3779B1u) >> (32 - kHashBits);
}
__device__ __forceinline__ uint32_t lanemask_lt() { uint32_t m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }

// Forward match length from `len0`, whole warp: 4 bytes per lane and step (match_kit/match_fast.rs:22-49).
__device__ __forceinline__ uint32_t warp_match_inc(const uint8_t *src, uint32_t a, uint32_t b, uint32_t len0, uint32_t max, uint32_t lane) {
    uint32_t len = len0;
    while (len < max) {
        const uint32_t off = len + lane * 4;
        uint32_t x = 0, valid = 0;  // valid = bytes of this lane's word that lie below max
        if (off < max) {
            valid = max - off < 4 ? max - off : 4;
            uint32_t wa = 0, wb = 0;
            if (valid == 4) { wa = ld_u32(src + a + off); wb = ld_u32(src + b + off); }
            else for (uint32_t k = 0; k < valid; k++) { wa |= (uint32_t)src[a + off + k] << (8 * k); wb |= (uint32_t)src[b + off + k] << (8 * k); }
            x = wa ^ wb;
        }
        const uint32_t diff = __ballot_sync(0xFFFFFFFFu, x != 0);
        if (diff) {
            const int l = __ffs(diff) - 1;
            const uint32_t xl = __shfl_sync(0xFFFFFFFFu, x, l);
            return len + l * 4 + ((__ffs(xl) - 1) >> 3);
        }
        len += 128;
    }
    return max;
}
// Backward match length, whole warp, one byte per lane and step (match_kit/match_fast.rs:61-89).
__device__ __forceinline__ uint32_t warp_match_dec(const uint8_t *src, uint32_t a, uint32_t b, uint32_t max, uint32_t lane) {
    uint32_t len = 0;
    while (len < max) {
        const uint32_t k = len + lane;
        const bool differ = k < max && src[a - k - 1] != src[b - k - 1];
        const uint32_t d = __ballot_sync(0xFFFFFFFFu, differ);
        if (d) return len + (__ffs(d) - 1);
        len += 32;
    }
    return max;
}

struct Match { uint32_t idx, match_idx, match_len; };

// Per-stream back-end state, identical in every lane (the control flow is warp-uniform).  Kept small:
// the parse kernel is occupancy-bound, so everything only needed when a block closes is re-read then.
struct FseSink {   // fse/buffer.rs + fse/backend.rs:66-96
    uint2 *packs;            // stream's pack scratch
    uint8_t *lits;           // stream's literal scratch
    uint32_t n_packs_total, n_lits_total;    // appended so far (all blocks)
    uint32_t blk_pack0, blk_lit0;            // where the open block starts
    uint32_t n_match_bytes, match_distance;
    uint32_t n_blocks;
    uint32_t out_used;
};
struct SinkEnv {  // what emit_block needs besides the sink (kernel parameters, re-read on use)
    const StreamCounts *bases;
    EncBlock *blocks;
    uint32_t *block_ids;
    uint32_t *block_counter;
    uint32_t si;
};

__device__ __noinline__ void sink_emit_block(FseSink &s, const SinkEnv &env, uint32_t lane) {  // emit_block_v2: close the open block
    if (lane == 0) {
        const StreamCounts base = env.bases[env.si];  // {packs, literal bytes, block slots, out bytes}
        EncBlock b;
        b.pack_off = base.n_blocks + s.blk_pack0;
        b.lit_off = base.n_fse + s.blk_lit0;
        b.out_off = base.n_lmds + s.out_used;
        b.n_packs = s.n_packs_total - s.blk_pack0;
        b.n_lits = s.n_lits_total - s.blk_lit0;
        b.n_match_bytes = s.n_match_bytes;
        b.out_size = 0;
        b.src_pos = 0; b.gather = 0; b.pad = 0; b.dst_pos = 0;
        const uint32_t id = atomicAdd(env.block_counter, 1u);
        env.blocks[id] = b;
        env.block_ids[base.n_literals + s.n_blocks] = id;
    }
    s.out_used += (uint32_t)((block_bound(s.n_lits_total - s.blk_lit0, s.n_packs_total - s.blk_pack0) + 15) & ~15ull);
    s.n_blocks++;
    s.blk_pack0 = s.n_packs_total; s.blk_lit0 = s.n_lits_total;
    s.n_match_bytes = 0; s.match_distance = 0;  // Buffer::reset
}
__device__ __forceinline__ void sink_push_pack(FseSink &s, uint32_t l, uint32_t m, uint32_t d, uint32_t lane) {
    if (lane == 0) s.packs[s.n_packs_total] = make_uint2(l | (m << 16), d);
    s.n_packs_total++;
}
__device__ __forceinline__ void sink_push_l(FseSink &s, uint32_t l, uint32_t lane) {  // Buffer::push_l
    s.match_distance = 1;
    sink_push_pack(s, l, 0, 1, lane);
}
__device__ __forceinline__ void sink_push_lmd(FseSink &s, uint32_t l, uint32_t m, uint32_t d, uint32_t lane) {  // Buffer::push_lmd
    if (s.match_distance == d) d = 0; else s.match_distance = d;
    sink_push_pack(s, l, m, d, lane);
    s.n_match_bytes += m;
}
__device__ __forceinline__ void sink_copy_lits(FseSink &s, const uint8_t *src, uint32_t from, uint32_t n, uint32_t lane) {
    uint8_t *dst = s.lits + s.n_lits_total;
    for (uint32_t t = lane; t < n; t += 32) dst[t] = src[from + t];
    s.n_lits_total += n;
}
// Buffer::push (fse/buffer.rs:45-97); returns false when the block is full and must be emitted first.
__device__ bool sink_buffer_push(FseSink &s, const uint8_t *src, uint32_t &lit_from, uint32_t &lit_len, uint32_t &match_len, uint32_t d, uint32_t lane) {
    while (lit_len > kMaxLValue) {
        if (s.n_packs_total - s.blk_pack0 == kLmdsPerBlock) return false;
        const uint32_t limit = kLiteralsPerBlock - (s.n_lits_total - s.blk_lit0);
        if (kMaxLValue <= limit) { sink_copy_lits(s, src, lit_from, kMaxLValue, lane); lit_from += kMaxLValue; lit_len -= kMaxLValue; sink_push_l(s, kMaxLValue, lane); }
        else if (limit != 0) { sink_copy_lits(s, src, lit_from, limit, lane); lit_from += limit; lit_len -= limit; sink_push_l(s, limit, lane); return false; }
        else return false;
    }
    if (s.n_packs_total - s.blk_pack0 == kLmdsPerBlock) return false;
    uint32_t literal_len = lit_len;
    const uint32_t limit = kLiteralsPerBlock - (s.n_lits_total - s.blk_lit0);
    if (literal_len <= limit) { sink_copy_lits(s, src, lit_from, literal_len, lane); lit_from += literal_len; lit_len = 0; }
    else if (limit != 0) { sink_copy_lits(s, src, lit_from, limit, lane); lit_from += limit; lit_len -= limit; sink_push_l(s, limit, lane); return false; }
    else return false;
    while (match_len > kMaxMValue) {
        sink_push_lmd(s, literal_len, kMaxMValue, d, lane);
        match_len -= kMaxMValue; literal_len = 0;
        if (s.n_packs_total - s.blk_pack0 == kLmdsPerBlock) return false;
    }
    sink_push_lmd(s, literal_len, match_len, d, lane);
    match_len = 0;
    return true;
}
__device__ void sink_push_match(FseSink &s, const SinkEnv &env, const uint8_t *src, uint32_t lit_from, uint32_t lit_len, uint32_t match_len, uint32_t d, uint32_t lane) {
    // common case of Buffer::push: one pack, block has room
    if (lit_len <= kMaxLValue && match_len <= kMaxMValue && s.n_packs_total - s.blk_pack0 < kLmdsPerBlock &&
        s.n_lits_total - s.blk_lit0 + lit_len <= kLiteralsPerBlock) {
        sink_copy_lits(s, src, lit_from, lit_len, lane);
        sink_push_lmd(s, lit_len, match_len, d, lane);
        return;
    }
    while (!sink_buffer_push(s, src, lit_from, lit_len, match_len, d, lane)) sink_emit_block(s, env, lane);  // fse/backend.rs:76-90
}

// LZVN back end (vn/backend.rs:37-136 + vn/opc.rs).  Lane 0 writes the opcode bytes.
struct VnSink {
    uint8_t *out;   // stream's out scratch; the 12-byte header is patched at the end
    uint32_t pos;   // bytes written (header included)
    uint32_t match_distance, n_literals, n_match_bytes;
};
__device__ __forceinline__ void vn_put(VnSink &v, uint32_t opu, uint32_t oplen, const uint8_t *src, uint32_t from, uint32_t n, uint32_t lane) {
    if (lane == 0) {
        for (uint32_t k = 0; k < oplen; k++) v.out[v.pos + k] = (uint8_t)(opu >> (8 * k));
        for (uint32_t k = 0; k < n; k++) v.out[v.pos + oplen + k] = src[from + k];
    }
    v.pos += oplen + n;
}
__device__ void vn_literal_runs(VnSink &v, const uint8_t *src, uint32_t &from, uint32_t &len, uint32_t keep_below, uint32_t lane) {
    while (len >= 0x10) {
        const uint32_t n = len < 0x10F ? len : 0x10F;
        vn_put(v, 0xE0u | ((n - 0x10) << 8), 2, src, from, n, lane);
        from += n; len -= n;
    }
    if (len >= keep_below && len > 0) { vn_put(v, 0xE0u | len, 1, src, from, len, lane); from += len; len = 0; }
}
__device__ void vn_push_match(VnSink &v, const uint8_t *src, uint32_t from, uint32_t lit_len, uint32_t match_len, uint32_t d, uint32_t lane) {
    v.n_literals += lit_len; v.n_match_bytes += match_len;
    vn_literal_runs(v, src, from, lit_len, 4, lane);
    const uint32_t L = lit_len;
    uint32_t n = 0x0A - 2 * L; if (n > match_len) n = match_len;
    match_len -= n;
    if (d == v.match_distance) {
        if (L == 0) vn_put(v, 0xF0u | n, 1, src, from, 0, lane);                                  // SmlM
        else vn_put(v, 0x6u | ((n - 3) << 3) | (L << 6), 1, src, from, L, lane);                  // PreD
    } else if (d < 0x600) {
        vn_put(v, ((d >> 8) & 7) | ((n - 3) << 3) | (L << 6) | ((d & 0xFF) << 8), 2, src, from, L, lane);  // SmlD
    } else if (d >= 0x4000 || match_len == 0 || n + match_len > 0x22) {
        vn_put(v, 0x7u | ((n - 3) << 3) | (L << 6) | (d << 8), 3, src, from, L, lane);            // LrgD
    } else {
        const uint32_t m = n - 3;
        vn_put(v, ((m >> 2) & 7) | (L << 3) | (0x5u << 5) | ((m & 3) << 8) | (d << 10), 3, src, from, L, lane);  // MedD
    }
    v.match_distance = d;
    while (match_len > 0x0F) { const uint32_t lim = match_len < 0x10F ? match_len : 0x10F; vn_put(v, 0xF0u | ((lim - 0x10) << 8), 2, src, from, 0, lane); match_len -= lim; }
    if (match_len > 0) vn_put(v, 0xF0u | match_len, 1, src, from, 0, lane);
}

// A history table is private to one warp, hence to one SM: its buckets can live in that SM's L1 (default loads and
// stores, prefetch into L1) instead of being fetched from L2 at every step.  LZB_PARSE_L1=0 restores the L2-only path.
#ifndef LZB_PARSE_CTAS
#define LZB_PARSE_CTAS 7   // resident 4-warp CTAs per SM (register bound: 72 registers -> 7)
#endif
#ifndef LZB_PARSE_SINK_SMEM
#define LZB_PARSE_SINK_SMEM 1
#endif
#ifndef LZB_PARSE_L1
#define LZB_PARSE_L1 1
#endif
struct Bucket { uint4 idx, val; };
__device__ __forceinline__ Bucket bucket_load(const uint32_t *bk) {
    Bucket b;
#if LZB_PARSE_L1
    b.idx = *reinterpret_cast<const uint4 *>(bk);
    b.val = *reinterpret_cast<const uint4 *>(bk + 4);
#else
    b.idx = __ldcg(reinterpret_cast<const uint4 *>(bk));
    b.val = __ldcg(reinterpret_cast<const uint4 *>(bk + 4));
#endif
    return b;
}
__device__ __forceinline__ void bucket_store(uint32_t *bk, const Bucket &b) {
#if LZB_PARSE_L1
    *reinterpret_cast<uint4 *>(bk) = b.idx;
    *reinterpret_cast<uint4 *>(bk + 4) = b.val;
#else
    __stcg(reinterpret_cast<uint4 *>(bk), b.idx);
    __stcg(reinterpret_cast<uint4 *>(bk + 4), b.val);
#endif
}
__device__ __forceinline__ void bucket_prefetch(const uint32_t *bk) {
#if LZB_PARSE_L1
    asm volatile("prefetch.global.L1 [%0];" ::"l"(bk));
#else
    asm volatile("prefetch.global.L2 [%0];" ::"l"(bk));
#endif
}
// Bucket after pushing this step's same-bucket positions: `me` = the newest one (this lane), `lower` = its lower
// peer lanes, `old` = the bucket before the step (HistoryTable::push x n, newest first, encode/history.rs:119-131).
// Every lane of the warp must call (shuffles).
// Positions are stored biased by the stream's epoch `vb` (see k_enc_parse); base_pos and me_idx are already biased.
__device__ __forceinline__ Bucket bucket_merge(const Bucket &old, uint32_t me_idx, uint32_t me_val, uint32_t lower, uint32_t base_pos, uint32_t val) {
    uint32_t m = lower, cnt = 1;
    int b1 = 0, b2 = 0, b3 = 0;
    if (m) { b1 = 31 - __clz(m); m &= ~(1u << b1); cnt = 2; }
    if (m) { b2 = 31 - __clz(m); m &= ~(1u << b2); cnt = 3; }
    if (m) { b3 = 31 - __clz(m); cnt = 4; }
    const uint32_t v1 = __shfl_sync(0xFFFFFFFFu, val, b1), v2 = __shfl_sync(0xFFFFFFFFu, val, b2), v3 = __shfl_sync(0xFFFFFFFFu, val, b3);
    Bucket nw;
    nw.idx.x = me_idx;
    nw.idx.y = cnt > 1 ? base_pos + b1 : old.idx.x;
    nw.idx.z = cnt > 2 ? base_pos + b2 : (cnt == 2 ? old.idx.x : old.idx.y);
    nw.idx.w = cnt > 3 ? base_pos + b3 : (cnt == 3 ? old.idx.x : (cnt == 2 ? old.idx.y : old.idx.z));
    nw.val.x = me_val;
    nw.val.y = cnt > 1 ? v1 : old.val.x;
    nw.val.z = cnt > 2 ? v2 : (cnt == 2 ? old.val.x : old.val.y);
    nw.val.w = cnt > 3 ? v3 : (cnt == 3 ? old.val.x : (cnt == 2 ? old.val.y : old.val.z));
    return nw;
}

constexpr int kParseWarps = 4;
constexpr uint32_t kFwdCap = 44;  // per-lane forward extension stops here (>= GOOD_MATCH_LEN); longer matches are finished warp-wide
constexpr uint32_t kBwdCap = 8;   // per-lane backward extension precomputed up to here

__device__ __forceinline__ uint32_t ld4u(const uint8_t *p) {  // unaligned 4-byte load: two aligned words + funnel shift
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t r = (uint32_t)a & 3u;
    const uint32_t *q = reinterpret_cast<const uint32_t *>(a - r);
    const uint32_t lo = q[0], hi = r ? q[1] : 0u;
    return __funnelshift_r(lo, hi, r * 8);
}

__device__ __forceinline__ uint64_t ld8u(const uint8_t *p) {  // unaligned 8-byte load: three aligned words + two funnel shifts
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t r = (uint32_t)a & 3u;
    const uint32_t *q = reinterpret_cast<const uint32_t *>(a - r);
    const uint32_t w0 = q[0], w1 = q[1], w2 = r ? q[2] : 0u;
    return (uint64_t)__funnelshift_r(w0, w1, r * 8) | ((uint64_t)__funnelshift_r(w1, w2, r * 8) << 32);
}

// Ordered insert of positions [from, to) into the history table, 32 per step (HistoryTable::push x n,
// encode/history.rs:24-31,119-131): the newest position of each bucket writes that bucket once.
__device__ __forceinline__ void history_insert_range(uint32_t *table, const uint8_t *src, uint32_t from, uint32_t to, bool vn, uint32_t lane, uint32_t vb) {
    while (from < to) {
        const uint32_t p = from + lane;
        const bool act = p < to;
        const uint32_t val = act ? ld4u(src + p) : 0u;
        const uint32_t h = act ? hash_u(val, vn) : (0xFFFF0000u + lane);
        const uint32_t peers = __match_any_sync(0xFFFFFFFFu, h);
        const bool writer = act && (peers & ~((2u << lane) - 1u)) == 0;  // newest position of its bucket
        uint32_t *bk = table + (writer ? h : 0u) * kBucketWords;
        Bucket old;
        old.idx = old.val = make_uint4(0, 0, 0, 0);
        if (writer) old = bucket_load(bk);
        const Bucket nw = bucket_merge(old, p + vb, val, peers & lanemask_lt(), from + vb, val);
        if (writer) bucket_store(bk, nw);
        __syncwarp();
        from += 32;
    }
}

__global__ void __launch_bounds__(kParseWarps * 32, LZB_PARSE_CTAS)
k_enc_parse(const uint8_t *__restrict__ src_base, const uint64_t *__restrict__ src_off, const uint64_t *__restrict__ src_len, size_t n_streams,
            EncStream *streams, const StreamCounts *__restrict__ bases, uint32_t *tables /* one per warp slot */, uint32_t *epochs, uint2 *pack_scratch,
            uint8_t *lit_scratch, uint32_t *block_ids, EncBlock *blocks, uint32_t *block_counter, uint8_t *out_scratch, uint32_t *stream_counter) {
    const uint32_t lane = lane_id();
    const uint32_t warp_slot = blockIdx.x * kParseWarps + (threadIdx.x >> 5);
    uint32_t *table = tables + (size_t)warp_slot * kTableWords;
    // HistoryTable::reset (encode/history.rs:72-83) makes every entry fail the distance test.  Instead of rewriting
    // 256 KiB per stream, positions are stored biased by an epoch that grows by more than (stream length + maximum
    // distance) from one stream to the next: whatever an earlier stream left behind is then too far away by
    // construction, exactly like a reset entry.  The epoch survives across launches next to the table; a table
    // fresh from cudaMemset (all zero) is covered by starting at kMaxDValue + 1.
    uint32_t epoch = epochs[warp_slot];
    if (epoch < kMaxDValue + 1) epoch = kMaxDValue + 1;
    for (;;) {
        // persistent warps pull streams from a counter; each owns one history table
        uint32_t si = 0;
        if (lane == 0) si = atomicAdd(stream_counter, 1u);
        si = __shfl_sync(0xFFFFFFFFu, si, 0);
        if (si >= n_streams) break;
        const uint32_t kind = streams[si].kind;
        if (kind == SK_RAW || streams[si].fast) continue;
        const bool vn = kind == SK_VN;
        const uint8_t *src = src_base + src_off[si];
        asm volatile("" : "+l"(src));  // one pointer in registers; otherwise every access re-adds the kernel parameter and the offset
        const uint32_t len = (uint32_t)src_len[si];
        const uint32_t max_d = vn ? kVnMaxD : kMaxDValue;
        if (epoch > 0xFFFFFFFFu - len - (kMaxDValue + 1)) {  // 32-bit positions would wrap inside this stream: really reset
            for (uint32_t t = lane; t < (1u << kHashBits); t += 32) *reinterpret_cast<uint4 *>(table + t * kBucketWords) = make_uint4(0, 0, 0, 0);
            epoch = kMaxDValue + 1;
            __syncwarp();
        }
        const uint32_t vb = epoch;
        epoch += len + kMaxDValue + 1;

        // The back-end state is warp-uniform and only touched in the sequential part.  It lives in shared memory: on
        // the stack (its address goes to the out-of-line block emitter) it competed for L1 with the history buckets
        // and the source, and every push waited for local-memory loads that had been evicted.
#if LZB_PARSE_SINK_SMEM
        __shared__ FseSink s_fs[kParseWarps];
        __shared__ VnSink s_vs[kParseWarps];
        FseSink &fs = s_fs[threadIdx.x >> 5];
        VnSink &vs = s_vs[threadIdx.x >> 5];
#else
        FseSink fs;
        VnSink vs;
#endif
        SinkEnv env;
        env.bases = bases; env.blocks = blocks; env.block_ids = block_ids; env.block_counter = block_counter; env.si = si;
        {
            const StreamCounts base = bases[si];
            fs.packs = pack_scratch + base.n_blocks; fs.lits = lit_scratch + base.n_fse;
            fs.n_packs_total = 0; fs.n_lits_total = 0; fs.blk_pack0 = 0; fs.blk_lit0 = 0; fs.n_match_bytes = 0; fs.match_distance = 0;
            fs.n_blocks = 0; fs.out_used = 0;
            vs.out = out_scratch + base.n_lmds; vs.pos = kVnHeaderSize; vs.match_distance = 0; vs.n_literals = 0; vs.n_match_bytes = 0;
        }
        auto push_match = [&](uint32_t lit_from, uint32_t lit_len, uint32_t match_len, uint32_t d) {
            if (vn) vn_push_match(vs, src, lit_from, lit_len, match_len, d, lane);
            else sink_push_match(fs, env, src, lit_from, lit_len, match_len, d, lane);
        };

        // FrontendBytes::match_any (encode/frontend_bytes.rs:160-211), 32 positions per step.
        //
        // Every position below the one being examined has been pushed into the history exactly once and in
        // order (:185-207 visits, :336-344 sync_history), so what position p finds in its bucket does not
        // depend on which matches were emitted.  Each lane therefore runs find_match (:214-244) for its own
        // position -- candidates = same-bucket positions of lower lanes (newest first), then the table's
        // bucket as of the step's start -- and the warp replays the sequential part (backward extension,
        // Match::select, push) over those results.
        const uint32_t end = len - 3;
        uint32_t index = 0, literal_index = 0;
        Match pending = {0, 0, 0};
        bool done = false;
        while (!done) {
            const uint32_t b0 = index;
            const uint32_t nb = end - b0 < 32 ? end - b0 : 32;
            const uint32_t p = b0 + lane;
            const bool act = lane < nb;
            // ---- phase 1: per-lane find_match ----
            uint32_t val = 0, h = 0xFFFF0000u + lane;
            Bucket tq;
            tq.idx = make_uint4(kEmptyIdx, kEmptyIdx, kEmptyIdx, kEmptyIdx);
            tq.val = make_uint4(0, 0, 0, 0);
            if (act) {
                val = ld4u(src + p);
                h = hash_u(val, vn);
                tq = bucket_load(table + h * kBucketWords);
            }
            const uint32_t peers = __match_any_sync(0xFFFFFFFFu, h);
            uint32_t earlier = peers & lanemask_lt();
            // Every position of the step ends up in the history (visited or skipped), so push them now: the
            // newest position of each bucket writes it once (same merge as history_insert_range).
            {
                const bool writer = act && (peers & ~((2u << lane) - 1u)) == 0;
                const Bucket nw = bucket_merge(tq, p + vb, val, earlier, b0 + vb, val);
                if (writer) bucket_store(table + h * kBucketWords, nw);
            }
            uint32_t c[4], cv[4];
            {   // first the same-bucket positions of lower lanes (newest first), then the table's bucket
                int e0 = 0, e1 = 0, e2 = 0, e3 = 0;
                uint32_t e = 0;
                if (earlier) { e0 = 31 - __clz(earlier); earlier &= ~(1u << e0); e = 1; }
                if (earlier) { e1 = 31 - __clz(earlier); earlier &= ~(1u << e1); e = 2; }
                if (earlier) { e2 = 31 - __clz(earlier); earlier &= ~(1u << e2); e = 3; }
                if (earlier) { e3 = 31 - __clz(earlier); e = 4; }
                const uint32_t w0 = __shfl_sync(0xFFFFFFFFu, val, e0), w1 = __shfl_sync(0xFFFFFFFFu, val, e1), w2 = __shfl_sync(0xFFFFFFFFu, val, e2),
                               w3 = __shfl_sync(0xFFFFFFFFu, val, e3);
                const uint32_t t0 = tq.idx.x - vb, t1 = tq.idx.y - vb, t2 = tq.idx.z - vb, t3 = tq.idx.w - vb;  // un-bias (stale entries stay out of range)
                c[0] = e > 0 ? b0 + e0 : t0;
                c[1] = e > 1 ? b0 + e1 : (e == 1 ? t0 : t1);
                c[2] = e > 2 ? b0 + e2 : (e == 2 ? t0 : (e == 1 ? t1 : t2));
                c[3] = e > 3 ? b0 + e3 : (e == 3 ? t0 : (e == 2 ? t1 : (e == 1 ? t2 : t3)));
                cv[0] = e > 0 ? w0 : tq.val.x;
                cv[1] = e > 1 ? w1 : (e == 1 ? tq.val.x : tq.val.y);
                cv[2] = e > 2 ? w2 : (e == 2 ? tq.val.x : (e == 1 ? tq.val.y : tq.val.z));
                cv[3] = e > 3 ? w3 : (e == 3 ? tq.val.x : (e == 2 ? tq.val.y : (e == 1 ? tq.val.z : tq.val.w)));
            }
            uint32_t r_len = 0, r_idx = 0, r_bw = 0;
            bool r_exact = false;  // best candidate hit the per-lane cap: lengths must be redone warp-wide
            if (act) {
                const uint32_t max = len - p;
                // Candidates are examined newest first up to the first one out of range (that also ends at the reset
                // marker); their first four bytes come with the bucket.
                const bool ok0 = p - c[0] <= max_d, ok1 = ok0 && p - c[1] <= max_d, ok2 = ok1 && p - c[2] <= max_d, ok3 = ok2 && p - c[3] <= max_d;
                const bool ok[4] = {ok0, ok1, ok2, ok3};
                // The extensions below run one candidate after the other and each starts with a load from a random
                // place of the stream (30 % of the kernel's stall samples sat on those four loads, r1c profile): ask
                // for the lines of all candidates that pass the 4-byte test first, so that the round trips overlap.
                // (Asking earlier, right after the bucket load, measured the same.)
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if (ok[k] && val == cv[k]) asm volatile("prefetch.global.L1 [%0];" ::"l"(src + c[k] + 4));
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    if (!ok[k]) continue;
                    const uint32_t x = val ^ cv[k];
                    uint32_t l = 0;
                    if (x == 0) {
                        l = 4;
                        while (l + 8 <= max && l < kFwdCap) {
                            const uint64_t y = ld8u(src + p + l) ^ ld8u(src + c[k] + l);
                            if (y) { l += (__ffsll((long long)y) - 1) >> 3; goto ext_done; }
                            l += 8;
                        }
                        if (l < kFwdCap) while (l < max && src[p + l] == src[c[k] + l]) l++;
                    ext_done:;
                    } else if (vn && (x & 0x00FFFFFFu) == 0) l = 3;
                    if (l >= kFwdCap && l < max) r_exact = true;
                    if (l > r_len) { r_len = l; r_idx = c[k]; }
                }
                if (r_len && !r_exact) {
                    const uint32_t lim = p < kBwdCap ? p : kBwdCap, lim2 = r_idx < lim ? r_idx : lim;
                    while (r_bw < lim2 && src[p - r_bw - 1] == src[r_idx - r_bw - 1]) r_bw++;
                }
            }
            // While this step is replayed, pull the next step's buckets towards the SM (hint only: the
            // history may still change before they are read).
            if (b0 + 32 + lane < end)
                bucket_prefetch(table + hash_u(ld4u(src + b0 + 32 + lane), vn) * kBucketWords);
            // ---- phase 2: the sequential front end over the step's positions ----
            // Positions whose find_match came back empty only advance the index (:203-209), so the replay
            // jumps from one position with a candidate to the next.
            const uint32_t has = __ballot_sync(0xFFFFFFFFu, act && r_len != 0);
            const uint32_t packed = r_len | (r_bw << 8) | ((uint32_t)r_exact << 16);  // r_len < kFwdCap + 4
            uint32_t cur = index;
            for (;;) {
                const uint32_t rel0 = cur - b0;
                const uint32_t m = rel0 < 32 ? (has & (0xFFFFFFFFu << rel0)) : 0u;
                if (m == 0) {  // nothing left in this step
                    if (cur < b0 + nb) cur = b0 + nb;
                    if (cur == end) done = true;
                    break;
                }
                const int l = __ffs(m) - 1;
                cur = b0 + l;
                const uint32_t pk = __shfl_sync(0xFFFFFFFFu, packed, l);
                Match inc;
                inc.match_len = pk & 0xFF;
                inc.match_idx = __shfl_sync(0xFFFFFFFFu, r_idx, l);
                inc.idx = cur;
                const bool exact = (pk >> 16) & 1;
                if (exact) {  // redo find_match for this position with the whole warp (uncapped lengths)
                    inc.match_len = 0; inc.match_idx = 0;
                    const uint32_t v = __shfl_sync(0xFFFFFFFFu, val, l);
                    for (int k = 0; k < 4; k++) {
                        const uint32_t cand = __shfl_sync(0xFFFFFFFFu, c[k], l);
                        if (cur - cand > max_d) break;
                        const uint32_t x = v ^ ld4u(src + cand);
                        uint32_t ml = 0;
                        if (x == 0) ml = warp_match_inc(src, cur, cand, 4, len - cur, lane);
                        else if (vn && (x & 0x00FFFFFFu) == 0) ml = 3;
                        if (ml > inc.match_len) { inc.match_len = ml; inc.match_idx = cand; }
                    }
                }
                {   // match_dec (:261-268)
                    const uint32_t lit = cur - literal_index;
                    const uint32_t lim = lit < inc.match_idx ? lit : inc.match_idx;
                    const uint32_t bw = (pk >> 8) & 0xFF;
                    uint32_t dec;
                    if (!exact && (bw < kBwdCap || lim <= kBwdCap)) dec = bw < lim ? bw : lim;
                    else dec = warp_match_dec(src, inc.idx, inc.match_idx, lim, lane);
                    inc.idx -= dec; inc.match_idx -= dec; inc.match_len += dec;
                }
                // Match::select (encode/match_object.rs:12-33), incoming.match_len != 0
                bool have = true;
                Match sel = pending;
                if (inc.match_len >= kGoodMatchLen) { sel = inc; pending.match_len = 0; }
                else if (pending.match_len == 0) { pending = inc; have = false; }
                else if ((int32_t)(pending.idx + pending.match_len - inc.idx) <= 0) { pending = inc; }
                else if (inc.match_len > pending.match_len) { sel = inc; pending.match_len = 0; }
                else { pending.match_len = 0; }
                if (have) {
                    push_match(literal_index, sel.idx - literal_index, sel.match_len, sel.idx - sel.match_idx);  // :287-302
                    literal_index = sel.idx + sel.match_len;
                    if (literal_index >= end) { done = true; break; }
                    cur = cur + 1 > literal_index ? cur + 1 : literal_index;  // index += 1; sync_history
                    if (cur >= end) { done = true; break; }
                } else {
                    cur++;
                    if (cur == end) { done = true; break; }
                }
            }
            // ---- phase 3: push every position this step passed into the history ----
            if (!done) {
                __syncwarp();  // the step's own positions were pushed in phase 1; a match may have run past them
                if (cur > b0 + nb) history_insert_range(table, src, b0 + nb, cur, vn, lane, vb);
            }
            index = cur;
        }
        // flush_pending, flush_literals, backend.finalize (:121-131,271-317)
        if (pending.match_len != 0) {
            push_match(literal_index, pending.idx - literal_index, pending.match_len, pending.idx - pending.match_idx);
            literal_index = pending.idx + pending.match_len;
        }
        if (!vn) {
            if (len - literal_index != 0) sink_push_match(fs, env, src, literal_index, len - literal_index, 0, 1, lane);  // push_literals
            sink_emit_block(fs, env, lane);  // finalize
            if (lane == 0) streams[si].n_blocks = fs.n_blocks;
        } else {
            uint32_t from = literal_index, ll = len - literal_index;
            if (ll != 0) { vs.n_literals += ll; vn_literal_runs(vs, src, from, ll, 0, lane); }
            vn_put(vs, 0x06u, 4, src, 0, 0, lane);   // 8-byte EOS: 06 00 00 00 00 00 00 00
            vn_put(vs, 0x00u, 4, src, 0, 0, lane);
            if (lane == 0) {
                uint8_t *h = vs.out;
                const uint32_t f[3] = {kMagicVxn, vs.n_literals + vs.n_match_bytes, vs.pos - kVnHeaderSize};
                for (int k = 0; k < 12; k++) h[k] = (uint8_t)(f[k >> 2] >> (8 * (k & 3)));
                streams[si].vn_size = vs.pos;
            }
        }
        __syncwarp();
    }
    if (lane == 0) epochs[warp_slot] = epoch;
}

// ------------------------------------------------------------------------------------------------
// Fast parse for bvx2 streams of <= kFastMaxLen bytes: k_enc_find (CTA per stream) + k_enc_replay (thread per stream).
//
// What a position finds in the history does not depend on the parse (see the top of this file), so find_match
// (encode/frontend_bytes.rs:214-244) can be evaluated for EVERY position of a stream in parallel, before the sequential
// front end runs.  The reference's HistoryTable (encode/history.rs:24-31,101-131: 2^14 buckets x the 4 newest positions)
// is replaced by the equivalent hash chain: prev[p] = the newest position q < p in p's bucket; the bucket as position p
// sees it is prev[p], prev[prev[p]], ... (4 hops).  With 16-bit positions the chain (128 KiB), the bucket heads
// (32 KiB) and the stream itself (64 KiB) fit one SM's shared memory, so every probe, compare and extension is a
// shared-memory access instead of a random HBM sector into a 512 KiB table per warp (r1: 242 GB of DRAM traffic per GiB).
//
//   k_enc_find    CTA (32 warps) per stream.  The stream is staged with one bulk async copy (cp.async.bulk + mbarrier).
//                 Warp 0 builds the chain in order, 32 positions per step (__match_any_sync resolves the step's own
//                 bucket collisions); warps 1..31 follow it unit by unit and compute, per position, the best candidate
//                 (strictly longest, newest first), its exact forward length and its backward length, packed into one
//                 32-bit word per position in HBM: distance[0:18] | min(len, 1023)[18:28] | min(bw, 15)[28:32].
//   k_enc_replay  thread per stream: the sequential front end (backward extension limit, Match::select, Buffer::push
//                 with splits and block closing) over those words; emits packs and block records only -- the literal
//                 bytes are collected by k_enc_fse_blocks from the packs.
// ------------------------------------------------------------------------------------------------
#ifndef LZB_CHAIN_MATES
#define LZB_CHAIN_MATES 7   // find warps allowed on the chain warp's scheduler (0..7)
#endif
#ifndef LZB_PRE_AHEAD
#define LZB_PRE_AHEAD 256   // info units handed out before the first find unit
#endif
#ifndef LZB_FENCE_SC
#define LZB_FENCE_SC 0
#endif
#ifndef LZB_POLL_NS
#define LZB_POLL_NS 100
#endif
#ifndef LZB_PRE_MATCH
#define LZB_PRE_MATCH 2   // 0: 14 ballots, 1: match.any, 2: 5 ballots + shuffle compares (measured 114K / 130K / 90K cycles per stream)
#endif
constexpr int kFindThreads = 1024;
constexpr uint32_t kFindUnit = 256;       // positions a find warp takes at a time
constexpr uint32_t kNone = 0xFFFFu;       // chain end (positions are < kFastMaxLen - 3)
constexpr uint32_t kLaneCap = 64;         // per-lane forward extension stops here; longer ones are finished warp-wide (or inherited from a run already measured)
constexpr uint32_t kWordLenSat = 1023, kWordBwSat = 15;
constexpr uint32_t kRunProbeAfter = 4096;  // cache hits after which a stream's positions ask the run cache before comparing bytes
constexpr uint32_t kRunSlots = 64;        // (distance, start, end) of long runs already measured, per stream
constexpr uint32_t kFsSrc = 16, kFsSrcBytes = kFastMaxLen + 48;  // 16 bytes below the stream (backward reads), up to 15 of alignment, over-read slack behind
constexpr uint32_t kFsHead = kFsSrc + kFsSrcBytes + 16, kFsPrev = kFsHead + (1u << kHashBits) * 2, kFsCtrl = kFsPrev + kFastMaxLen * 2;
struct FindCtrl {
    unsigned long long mbar;
    uint32_t chain_done, next_ticket, stream;
    uint32_t run_hits;  // positions of this stream that found their length in the run cache (saturates at kRunProbeAfter)
    unsigned long long runs[kRunSlots];
    uint8_t pre_done[kFastMaxLen / kFindUnit];  // per unit: peer info written
};
constexpr uint32_t kFindSmemBytes = kFsCtrl + sizeof(FindCtrl);
static_assert(kFindSmemBytes <= 232448, "k_enc_find shared memory");

__device__ __forceinline__ uint32_t lds_u32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t lds_u16(uint32_t a) { uint32_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t lds_b8(uint32_t a) { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts_u16(uint32_t a, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
// unaligned reads from shared memory: aligned words + funnel shifts
__device__ __forceinline__ uint32_t lds4u(uint32_t a) {
    const uint32_t r = a & 3u, q = a - r;
    return __funnelshift_r(lds_u32(q), lds_u32(q + 4), r * 8);
}
__device__ __forceinline__ uint64_t lds8u(uint32_t a) {
    const uint32_t r = a & 3u, q = a - r;
    const uint32_t w0 = lds_u32(q), w1 = lds_u32(q + 4), w2 = lds_u32(q + 8);
    return (uint64_t)__funnelshift_r(w0, w1, r * 8) | ((uint64_t)__funnelshift_r(w1, w2, r * 8) << 32);
}
// Forward match length of s[a..] against s[b..] starting from `l` known equal bytes, at most `lim` (match_kit/match_fast.rs:22-49).
__device__ __forceinline__ uint32_t smem_match_inc(uint32_t s, uint32_t a, uint32_t b, uint32_t l, uint32_t lim) {
    while (l + 8 <= lim) {
        const uint64_t y = lds8u(s + a + l) ^ lds8u(s + b + l);
        if (y) return l + ((__ffsll((long long)y) - 1) >> 3);
        l += 8;
    }
    while (l < lim && lds_b8(s + a + l) == lds_b8(s + b + l)) l++;
    return l;
}
// The same, whole warp, 8 bytes per lane and step (for runs longer than kLaneCap).
__device__ __forceinline__ uint32_t smem_warp_match_inc(uint32_t s, uint32_t a, uint32_t b, uint32_t l, uint32_t lim, uint32_t lane) {
    while (l < lim) {
        const uint32_t off = l + lane * 8;
        uint64_t y = 0;
        if (off + 8 <= lim) y = lds8u(s + a + off) ^ lds8u(s + b + off);
        else for (uint32_t k = 0; off + k < lim; k++) y |= (uint64_t)(lds_b8(s + a + off + k) ^ lds_b8(s + b + off + k)) << (8 * k);
        const uint32_t diff = __ballot_sync(0xFFFFFFFFu, y != 0);
        if (diff) {
            const int j = __ffs(diff) - 1;
            const uint64_t yj = __shfl_sync(0xFFFFFFFFu, y, j);
            return l + j * 8 + ((__ffsll((long long)yj) - 1) >> 3);
        }
        l += 256;
    }
    return lim;
}
// release / acquire fence at CTA scope around the shared-memory flags (MEMBAR.SC.CTA, what __threadfence_block() emits, is heavier)
__device__ __forceinline__ void fence_cta() {
#if LZB_FENCE_SC
    __threadfence_block();
#else
    asm volatile("fence.acq_rel.cta;" ::: "memory");
#endif
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t phase) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(mbar), "r"(phase) : "memory");
}

// find_match's candidate loop for a position of a PERIODIC stream (see k_enc_find): the run cache is asked before any byte is
// compared; a position well inside a run that was already measured at this distance knows its length without looking at the
// bytes.  res = {best_len, best_c, n_sat, cs[4], ls[4]}.
__device__ __noinline__ void find_candidates_probe(uint32_t s, uint32_t s_prev, const FindCtrl *ctrl, uint32_t p, uint32_t maxl, uint32_t *res) {
    uint32_t best_len = 0, best_c = 0, n_sat = 0;
    uint32_t cs[4] = {0, 0, 0, 0}, ls[4] = {0, 0, 0, 0};
    const uint32_t val = lds4u(s + p);
    uint32_t c = lds_u16(s_prev + p * 2);
    if (c != kNone) {
        const uint64_t p8 = lds8u(s + p + 4);
        const uint32_t lim = maxl < kLaneCap ? maxl : kLaneCap;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t cn = lds_u16(s_prev + c * 2);
            if (lds4u(s + c) == val) {
                uint32_t l;
                const uint32_t d = p - c;
                const unsigned long long e = *reinterpret_cast<const volatile unsigned long long *>(&ctrl->runs[(d * 0x9E3779B1u) >> 26]);
                const uint32_t ed = (uint32_t)(e & 0x3FFFF), es = (uint32_t)(e >> 18) & 0x1FFFFF, ee = (uint32_t)(e >> 39);
                if (ed == d && es <= p && p + kLaneCap < ee) {
                    l = kLaneCap; cs[n_sat] = c; ls[n_sat] = ee - p; n_sat++;
                } else {
                    const uint64_t y = p8 ^ lds8u(s + c + 4);
                    if (y) { l = 4 + ((__ffsll((long long)y) - 1) >> 3); l = l < lim ? l : lim; }
                    else l = lim >= 12 ? smem_match_inc(s, p, c, 12, lim) : lim;
                    if (l == kLaneCap && l < maxl) { cs[n_sat] = c; ls[n_sat] = 0; n_sat++; }
                }
                if (l > best_len) { best_len = l; best_c = c; }
            }
            c = cn;
            if (c == kNone) break;
        }
    }
    res[0] = best_len; res[1] = best_c; res[2] = n_sat;
#pragma unroll
    for (int k = 0; k < 4; k++) { res[3 + k] = cs[k]; res[7 + k] = ls[k]; }
}

__global__ void __launch_bounds__(kFindThreads, 1)
k_enc_find(const uint8_t *__restrict__ src_base, const uint64_t *__restrict__ src_off, const uint64_t *__restrict__ src_len, size_t n_streams,
           const EncStream *__restrict__ streams, const StreamCounts *__restrict__ bases, uint32_t *__restrict__ words, uint32_t *stream_counter) {
    extern __shared__ __align__(128) uint8_t fsm[];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t sm0 = (uint32_t)__cvta_generic_to_shared(fsm);
    FindCtrl *ctrl = reinterpret_cast<FindCtrl *>(fsm + kFsCtrl);
    const uint32_t mbar = sm0 + kFsCtrl;
    const uint32_t s_head = sm0 + kFsHead, s_prev = sm0 + kFsPrev;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t phase = 0;
    for (;;) {
        if (tid == 0) ctrl->stream = atomicAdd(stream_counter, 1u);
        __syncthreads();
        const uint32_t si = ctrl->stream;
        if (si >= n_streams) break;
        if (streams[si].fast != 1) { __syncthreads(); continue; }
        const uint32_t len = (uint32_t)src_len[si], end = len - 3;
        const uint32_t n_units = (end + kFindUnit - 1) / kFindUnit;
        const uint8_t *g = src_base + src_off[si];
        const uint32_t mis = (uint32_t)reinterpret_cast<uintptr_t>(g) & 15u;
        const uint32_t s = sm0 + kFsSrc + mis;  // shared address of stream byte 0
        if (tid == 0) {
            // One bulk async copy stages the whole stream: 16-byte aligned on both sides, the stream's misalignment kept.
            // (The last 16-byte unit may reach past the stream's end; it lies inside the same 16-byte aligned granule as
            // the last stream byte, hence inside the caller's allocation.)
            const uint32_t bytes = (mis + len + 15u) & ~15u;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sm0 + kFsSrc),
                         "l"(g - mis), "r"(bytes), "r"(mbar)
                         : "memory");
            ctrl->chain_done = 0; ctrl->next_ticket = 0; ctrl->run_hits = 0;
        }
        // meanwhile: empty bucket heads, run cache, unit flags
        for (uint32_t t = tid; t < (1u << kHashBits) * 2 / 16; t += kFindThreads)
            *reinterpret_cast<uint4 *>(fsm + kFsHead + t * 16) = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
        if (tid < kRunSlots) ctrl->runs[tid] = 0;
        if (tid < kFastMaxLen / kFindUnit / 4) reinterpret_cast<uint32_t *>(ctrl->pre_done)[tid] = 0;
        mbar_wait(mbar, phase);
        phase ^= 1;
        __syncthreads();
        uint32_t *wout = words + bases[si].n_fse;
#ifdef LZB_FIND_DEBUG
        const long long dbg_t0 = clock64();
        long long dbg_chain = 0, dbg_pre = 0, dbg_find0 = 0;
        bool dbg_passed = false;
#endif
        if (warp == 0) {
            // ---- the chain: HistoryTable::push for every position, in order, 32 per step ----
            // The bucket index of every position and which lanes of a step share a bucket were worked out by the other
            // warps (info word parked in prev[p], see below), so a step is: bucket head -> prev[p]; newest position of each
            // bucket -> head.  One warp runs this serial chain; what it costs is instructions on its dependent path
            // (a lone warp issues one every ~4 cycles), hence everything that can be precomputed is.
            // Only the bucket heads carry a dependency from one step to the next, and that one is satisfied by program order
            // (a step's head stores are issued before the next step's head loads); nothing in the chain consumes the loaded
            // values.  So a unit's eight steps issue their loads and stores back to back and prev[] is written at the end:
            // one shared-memory round trip per unit instead of one per step.
            for (uint32_t u = 0; u < n_units; u++) {
                while (*reinterpret_cast<volatile uint8_t *>(&ctrl->pre_done[u]) == 0) {}
                fence_cta();
                const uint32_t ub = u * kFindUnit;
                constexpr uint32_t kSteps = kFindUnit / 32;
                uint32_t info[kSteps], h[kSteps], old[kSteps];
#pragma unroll
                for (uint32_t k = 0; k < kSteps; k++) info[k] = lds_u16(s_prev + (ub + k * 32 + lane) * 2);  // stays inside prev[]: ub + 255 <= 65535
#pragma unroll
                for (uint32_t k = 0; k < kSteps; k++) {
                    const uint32_t hb = __shfl_sync(0xFFFFFFFFu, info[k], (info[k] >> 5) & 31u);  // the bucket index lives in the group's lowest lane
                    h[k] = ((info[k] & 0x8000u) ? hb : info[k]) & 0x3FFFu;
                }
#pragma unroll
                for (uint32_t k = 0; k < kSteps; k++) {
                    const uint32_t p = ub + k * 32 + lane;
                    old[k] = lds_u16(s_head + h[k] * 2);
                    if ((info[k] & 0x4000u) && p < end) sts_u16(s_head + h[k] * 2, p);  // newest position of its bucket in this step
                    __syncwarp();
                }
#pragma unroll
                for (uint32_t k = 0; k < kSteps; k++) {
                    const uint32_t b0 = ub + k * 32, p = b0 + lane;
                    if (p < end) sts_u16(s_prev + p * 2, (info[k] & 0x8000u) ? b0 + (info[k] & 31u) : old[k]);
                }
                __syncwarp();
                fence_cta();
                if (lane == 0) *reinterpret_cast<volatile uint32_t *>(&ctrl->chain_done) = ub + kFindUnit < end ? ub + kFindUnit : end;
            }
#ifdef LZB_FIND_DEBUG
            dbg_chain = clock64() - dbg_t0;
#endif
#ifdef LZB_FIND_PHASED
            __syncthreads();  // chain done
#endif
        } else if ((warp & 3u) != 0 || (warp >> 2) <= LZB_CHAIN_MATES) {
            // (Warps that would share the chain warp's scheduler beyond LZB_CHAIN_MATES of them stay idle: the chain is one
            // warp's dependent instruction stream, and every warp issuing next to it stretches it.)
            for (;;) {
                uint32_t t = 0;
                if (lane == 0) t = atomicAdd(&ctrl->next_ticket, 1u);
                t = __shfl_sync(0xFFFFFFFFu, t, 0);
                if (t >= 2 * n_units) break;
#ifdef LZB_FIND_PHASED
                if (t == n_units) {}  // handled below
#endif
                // ticket order: the info units run LZB_PRE_AHEAD units ahead of the find units, then alternate with them
                {
                    const uint32_t ahead = LZB_PRE_AHEAD < n_units ? LZB_PRE_AHEAD : n_units;
                    if (t < ahead) {}                                            // info unit t
                    else if (t < ahead + 2 * (n_units - ahead)) {                // alternate: find unit j, info unit ahead + j
                        const uint32_t j = (t - ahead) >> 1;
                        t = ((t - ahead) & 1u) ? ahead + j : n_units + j;
                    } else t = n_units + (t - ahead - (n_units - ahead));        // the last `ahead` find units
                }
#ifdef LZB_FIND_PHASED
                if (t >= n_units && !dbg_passed) {
                    dbg_passed = true;
                    dbg_pre = clock64() - dbg_t0;
                    __syncthreads();  // all info units done (every warp gets here: tickets >= n_units exist for all of them or the loop ends)
                    __syncthreads();  // chain done
                    dbg_find0 = clock64() - dbg_t0;
                }
#endif
                if (t < n_units) {
                    // ---- info for the chain: bucket index, and which lanes of each 32-position step share a bucket ----
                    // (__match_any_sync takes ~350 cycles when all 32 values differ, which they nearly always do; 14 ballots
                    // over the bits of the bucket index give the same mask.)  Parked in prev[p] until the chain gets there:
                    //   no lower lane in my bucket:  bit 14 = no higher lane either (I am the newest), bits 0..13 = bucket
                    //   otherwise: bit 15, bit 14 as above, bits 5..9 = lowest lane of the group, bits 0..4 = next lower lane
                    const uint32_t p0 = t * kFindUnit, p1 = p0 + kFindUnit < end ? p0 + kFindUnit : end;
                    for (uint32_t b0 = p0; b0 < p1; b0 += 32) {
                        const uint32_t p = b0 + lane;
                        const bool act = p < end;
                        const uint32_t h = hash_u(lds4u(s + (act ? p : end - 1)), false);
#if LZB_PRE_MATCH == 1
                        uint32_t m = __match_any_sync(0xFFFFFFFFu, act ? h : 0xFFFF0000u + lane);
#elif LZB_PRE_MATCH == 2
                        // five ballots split the lanes into classes by the low bits of the bucket index; the few lanes of a
                        // class are then compared one by one (shuffle) -- VOTE is the scarce pipe here
                        uint32_t mc = __ballot_sync(0xFFFFFFFFu, act);
#pragma unroll
                        for (int bit = 0; bit < 5; bit++) {
                            const bool on = (h >> bit) & 1u;
                            const uint32_t bl = __ballot_sync(0xFFFFFFFFu, on);
                            mc &= on ? bl : ~bl;
                        }
                        uint32_t m = act ? (1u << lane) : 0u;
                        uint32_t rest = mc & ~(1u << lane);
                        while (__any_sync(0xFFFFFFFFu, rest != 0)) {
                            const int j = rest ? __ffs(rest) - 1 : (int)lane;
                            const uint32_t hj = __shfl_sync(0xFFFFFFFFu, h, j);
                            if (rest && hj == h) m |= 1u << j;
                            rest &= rest - 1;
                        }
#else
                        uint32_t m = __ballot_sync(0xFFFFFFFFu, act);
#pragma unroll
                        for (int bit = 0; bit < (int)kHashBits; bit++) {
                            const bool on = (h >> bit) & 1u;
                            const uint32_t bl = __ballot_sync(0xFFFFFFFFu, on);
                            m &= on ? bl : ~bl;
                        }
#endif
                        const uint32_t lower = m & lanemask_lt();
                        const uint32_t newest = (m >> lane) == 1u ? 0x4000u : 0u;
                        const uint32_t info = lower ? (0x8000u | newest | ((uint32_t)(__ffs(m) - 1) << 5) | (31 - __clz(lower))) : (newest | h);
                        if (act) sts_u16(s_prev + p * 2, info);
                    }
                    __syncwarp();
                    fence_cta();
                    if (lane == 0) *reinterpret_cast<volatile uint8_t *>(&ctrl->pre_done[t]) = 1;
#ifdef LZB_FIND_DEBUG
                    dbg_pre = clock64() - dbg_t0;
#endif
                    continue;
                }
                // ---- find_match for every position of a unit, behind the chain ----
                const uint32_t p0 = (t - n_units) * kFindUnit;
                const uint32_t p1 = p0 + kFindUnit < end ? p0 + kFindUnit : end;
                while (*reinterpret_cast<volatile uint32_t *>(&ctrl->chain_done) < p1) __nanosleep(LZB_POLL_NS);
                fence_cta();
                for (uint32_t b0 = p0; b0 < p1; b0 += 32) {
                    const uint32_t p = b0 + lane;
                    const bool act = p < p1;
                    // Once a stream has shown itself to be periodic (its positions keep finding their lengths in the run cache),
                    // the cache is asked BEFORE the bytes are compared; on text that question would only cost (measured +14 %).
                    const uint32_t hits_seen = *reinterpret_cast<volatile uint32_t *>(&ctrl->run_hits);
                    const bool probe_first = hits_seen >= kRunProbeAfter;
                    uint32_t best_len = 0, best_c = 0;
                    uint32_t cs[4], ls[4];  // candidates whose length reached kLaneCap
                    uint32_t n_sat = 0;
                    const uint32_t maxl = len - p;
                    // (two copies of the candidate loop: the one that asks the run cache first is out of line and only entered by
                    // periodic streams, so that text runs exactly the loop it ran before)
                    if (probe_first) {
                        if (act) {
                            uint32_t res[11];
                            find_candidates_probe(s, s_prev, ctrl, p, maxl, res);
                            best_len = res[0]; best_c = res[1]; n_sat = res[2];
#pragma unroll
                            for (int k = 0; k < 4; k++) { cs[k] = res[3 + k]; ls[k] = res[7 + k]; }
                        }
                    } else if (act) {
                        const uint32_t val = lds4u(s + p);
                        uint32_t c = lds_u16(s_prev + p * 2);
                        if (c != kNone) {
                            const uint64_t p8 = lds8u(s + p + 4);  // bytes 4..11 of this position, shared by its candidates
                            const uint32_t lim = maxl < kLaneCap ? maxl : kLaneCap;
#pragma unroll
                            for (int k = 0; k < 4; k++) {
                                // (the distance limit 262 139 cannot be exceeded inside 64 KiB)
                                const uint32_t cn = lds_u16(s_prev + c * 2);
                                if (lds4u(s + c) == val) {
                                    const uint64_t y = p8 ^ lds8u(s + c + 4);
                                    uint32_t l;
                                    if (y) { l = 4 + ((__ffsll((long long)y) - 1) >> 3); l = l < lim ? l : lim; }
                                    else l = lim >= 12 ? smem_match_inc(s, p, c, 12, lim) : lim;
                                    if (l == kLaneCap && l < maxl) { cs[n_sat] = c; ls[n_sat] = 0; n_sat++; }
                                    if (l > best_len) { best_len = l; best_c = c; }
                                }
                                c = cn;
                                if (c == kNone) break;
                            }
                        }
                    }
                    // Candidates beyond the per-lane cap need their exact lengths (strictly longest wins, newest first, and the
                    // word stores up to 1023).  Runs are measured once per (distance, start) with the whole warp and remembered:
                    // inside a run every later position inherits end - p, which keeps periodic data linear.
                    bool need = n_sat >= 1;
                    while (__any_sync(0xFFFFFFFFu, need)) {
                        if (need) {
                            bool open = false;
                            for (uint32_t k = 0; k < n_sat; k++) {
                                if (ls[k]) continue;
                                const uint32_t d = p - cs[k];
                                const unsigned long long e = *reinterpret_cast<volatile unsigned long long *>(&ctrl->runs[(d * 0x9E3779B1u) >> 26]);
                                const uint32_t ed = (uint32_t)(e & 0x3FFFF), es = (uint32_t)(e >> 18) & 0x1FFFFF, ee = (uint32_t)(e >> 39);
                                if (ed == d && es <= p && p + kLaneCap <= ee) { ls[k] = ee - p; if (hits_seen < kRunProbeAfter) atomicAdd(&ctrl->run_hits, 1u); }
                                else open = true;
                            }
                            need = open;
                        }
                        const uint32_t todo = __ballot_sync(0xFFFFFFFFu, need);
                        if (!todo) break;
                        const int j = __ffs(todo) - 1;
                        uint32_t k0 = 0;
                        while (k0 < 3 && !(k0 < n_sat && ls[k0] == 0)) k0++;
                        const uint32_t pj = __shfl_sync(0xFFFFFFFFu, p, j), cj = __shfl_sync(0xFFFFFFFFu, cs[k0 & 3], j);
                        const uint32_t ej = pj + smem_warp_match_inc(s, pj, cj, kLaneCap, len - pj, lane);
                        if (lane == 0) {
                            const uint32_t d = pj - cj;
                            *reinterpret_cast<volatile unsigned long long *>(&ctrl->runs[(d * 0x9E3779B1u) >> 26]) =
                                (unsigned long long)d | ((unsigned long long)pj << 18) | ((unsigned long long)ej << 39);
                        }
                        if (lane == (uint32_t)j) ls[k0 & 3] = ej - pj;  // progress even if another warp reuses the slot
                        __syncwarp();
                        fence_cta();
                    }
                    uint32_t word = 0;
                    if (act && best_len) {
                        if (n_sat >= 1) {  // candidates are in newest-first order: strictly longer wins (any capped one beats the others)
                            uint32_t bl = 0;
                            for (uint32_t k = 0; k < n_sat; k++)
                                if (ls[k] > bl) { bl = ls[k]; best_c = cs[k]; }
                            best_len = bl;
                        }
                        // backward length (match_kit/match_fast.rs:61-89), 4 bytes per step, up to what the word can hold
                        uint32_t bw = 0;
                        const uint32_t blim = best_c < kWordBwSat ? best_c : kWordBwSat;
                        while (bw < blim) {
                            const uint32_t x = lds4u(s + p - bw - 4) ^ lds4u(s + best_c - bw - 4);  // reads at most 4 bytes below the stream (padding)
                            const uint32_t nb = x ? (uint32_t)__clz(x) >> 3 : 4u;
                            bw += nb;
                            if (nb < 4) break;
                        }
                        bw = bw < blim ? bw : blim;
                        word = (p - best_c) | ((best_len < kWordLenSat ? best_len : kWordLenSat) << 18) | (bw << 28);
                    }
                    // positions without a candidate carry the distance to the next position of this 32-group that has one
                    // (distance field 0, length field = skip), so the front end steps over empty runs in one go
                    {
                        const uint32_t nz = __ballot_sync(0xFFFFFFFFu, word != 0);
                        if (word == 0) {
                            const uint32_t next = lane == 31 ? 0u : nz & (0xFFFFFFFEu << lane);
                            word = (next ? (uint32_t)__ffs(next) - 1u - lane : 32u - lane) << 18;
                        }
                    }
                    if (act) wout[p] = word;
                }
            }
        }
#ifdef LZB_FIND_DEBUG
        const long long dbg_mine = clock64() - dbg_t0;
#ifdef LZB_FIND_PHASED
        if (warp != 0 && !dbg_passed) { __syncthreads(); __syncthreads(); }
        if (blockIdx.x == 3 && lane == 0 && warp == 1) printf("  phased: pre_end %lld find_start %lld\n", dbg_pre, dbg_find0);
#endif
        __syncthreads();
        if (blockIdx.x == 3 && lane == 0 && (warp < 3 || warp == 31)) printf("stream %u warp %u: chain %lld lastpre %lld mywork_end %lld all_end %lld\n", si, warp, dbg_chain, dbg_pre, dbg_mine, clock64() - dbg_t0);
#else
        __syncthreads();
#endif
    }
}

// The sequential front end, one THREAD per stream (32 streams per warp): FrontendBytes::match_any's control flow
// (encode/frontend_bytes.rs:160-211,261-317), Match::select (encode/match_object.rs:12-33), FseBackend::push_match /
// Buffer::push (fse/backend.rs:66-96, fse/buffer.rs:45-117) over the per-position words of k_enc_find.
struct TSink {
    uint2 *packs;
    uint32_t n_packs_total, n_lits_total, blk_pack0, blk_lit0, n_match_bytes, match_distance, n_blocks, out_used, blk_src0;
};
struct TEnv {
    StreamCounts base;
    EncBlock *blocks;
    uint32_t *block_ids;
    uint32_t *block_counter;
    uint64_t src_off;
};
__device__ __forceinline__ void tsink_emit_block(TSink &s, const TEnv &env) {
    EncBlock b;
    b.pack_off = env.base.n_blocks + s.blk_pack0;
    b.lit_off = env.base.n_fse + s.blk_lit0;
    b.out_off = env.base.n_lmds + s.out_used;
    b.src_pos = env.src_off + s.blk_src0;
    b.n_packs = s.n_packs_total - s.blk_pack0;
    b.n_lits = s.n_lits_total - s.blk_lit0;
    b.n_match_bytes = s.n_match_bytes;
    b.out_size = 0; b.gather = 1; b.pad = 0; b.dst_pos = 0;
    const uint32_t id = atomicAdd(env.block_counter, 1u);
    env.blocks[id] = b;
    env.block_ids[env.base.n_literals + s.n_blocks] = id;
    s.out_used += (uint32_t)((block_bound(b.n_lits, b.n_packs) + 15) & ~15ull);
    s.n_blocks++;
    s.blk_src0 += b.n_lits + b.n_match_bytes;
    s.blk_pack0 = s.n_packs_total; s.blk_lit0 = s.n_lits_total;
    s.n_match_bytes = 0; s.match_distance = 0;
}
__device__ __forceinline__ void tsink_push_l(TSink &s, uint32_t l) {
    s.match_distance = 1;
    s.packs[s.n_packs_total++] = make_uint2(l, 1);
}
__device__ __forceinline__ void tsink_push_lmd(TSink &s, uint32_t l, uint32_t m, uint32_t d) {
    if (s.match_distance == d) d = 0; else s.match_distance = d;
    s.packs[s.n_packs_total++] = make_uint2(l | (m << 16), d);
    s.n_match_bytes += m;
}
__device__ __forceinline__ bool tsink_buffer_push(TSink &s, uint32_t &lit_len, uint32_t &match_len, uint32_t d) {  // Buffer::push
    while (lit_len > kMaxLValue) {
        if (s.n_packs_total - s.blk_pack0 == kLmdsPerBlock) return false;
        const uint32_t limit = kLiteralsPerBlock - (s.n_lits_total - s.blk_lit0);
        if (kMaxLValue <= limit) { s.n_lits_total += kMaxLValue; lit_len -= kMaxLValue; tsink_push_l(s, kMaxLValue); }
        else if (limit != 0) { s.n_lits_total += limit; lit_len -= limit; tsink_push_l(s, limit); return false; }
        else return false;
    }
    if (s.n_packs_total - s.blk_pack0 == kLmdsPerBlock) return false;
    uint32_t literal_len = lit_len;
    const uint32_t limit = kLiteralsPerBlock - (s.n_lits_total - s.blk_lit0);
    if (literal_len <= limit) { s.n_lits_total += literal_len; lit_len = 0; }
    else if (limit != 0) { s.n_lits_total += limit; lit_len -= limit; tsink_push_l(s, limit); return false; }
    else return false;
    while (match_len > kMaxMValue) {
        tsink_push_lmd(s, literal_len, kMaxMValue, d);
        match_len -= kMaxMValue; literal_len = 0;
        if (s.n_packs_total - s.blk_pack0 == kLmdsPerBlock) return false;
    }
    tsink_push_lmd(s, literal_len, match_len, d);
    match_len = 0;
    return true;
}
__device__ __forceinline__ void tsink_push_match(TSink &s, const TEnv &env, uint32_t lit_len, uint32_t match_len, uint32_t d) {
    if (lit_len <= kMaxLValue && match_len <= kMaxMValue && s.n_packs_total - s.blk_pack0 < kLmdsPerBlock &&
        s.n_lits_total - s.blk_lit0 + lit_len <= kLiteralsPerBlock) {
        s.n_lits_total += lit_len;
        tsink_push_lmd(s, lit_len, match_len, d);
        return;
    }
    while (!tsink_buffer_push(s, lit_len, match_len, d)) tsink_emit_block(s, env);
}

#ifndef LZB_REPLAY_STRIDE
#define LZB_REPLAY_STRIDE 1   // measured 1/2/4/8/16: 11.1 / 11.4 / 12.4 / 15.4 / 23.8 ms -- fewer streams per warp do not help
#endif
constexpr int kReplayThreads = 32;
constexpr int kReplayStride = LZB_REPLAY_STRIDE;
#ifndef LZB_RING_WORDS
#define LZB_RING_WORDS 128   // 64: replay 14.6 ms per GiB (half a ring of lead time is about one memory round trip on text)
#endif
constexpr uint32_t kRingWords = LZB_RING_WORDS, kRingStride = kRingWords * 4 + 16;  // per-lane ring of words (stride skews the banks)
__global__ void __launch_bounds__(kReplayThreads)
k_enc_replay(const uint8_t *__restrict__ src_base, const uint64_t *__restrict__ src_off, const uint64_t *__restrict__ src_len, size_t n_streams,
             EncStream *streams, const StreamCounts *__restrict__ bases, const uint32_t *__restrict__ words, uint2 *pack_scratch, uint32_t *block_ids,
             EncBlock *blocks, uint32_t *block_counter) {
    // The 32 lanes of a warp (= one CTA) replay 32 different streams.  All of them stay in the loop until the last one
    // is done, because the ring upkeep below is WARP-WIDE on purpose (see there).
    static_assert(kReplayStride == 1, "one stream per lane");
    const size_t si = (size_t)blockIdx.x * kReplayThreads + threadIdx.x;
    const bool valid = si < n_streams && streams[si < n_streams ? si : 0].fast == 1 && !streams[si < n_streams ? si : 0].seg;
    bool active = valid;
    const size_t sj = valid ? si : 0;
    const uint8_t *src = src_base + src_off[sj];
    const uint32_t len = valid ? (uint32_t)src_len[sj] : 4u, end = len - 3;
    TEnv env;
    env.base = bases[sj]; env.blocks = blocks; env.block_ids = block_ids; env.block_counter = block_counter; env.src_off = src_off[sj];
    TSink fs;
    fs.packs = pack_scratch + env.base.n_blocks;
    fs.n_packs_total = 0; fs.n_lits_total = 0; fs.blk_pack0 = 0; fs.blk_lit0 = 0; fs.n_match_bytes = 0; fs.match_distance = 0;
    fs.n_blocks = 0; fs.out_used = 0; fs.blk_src0 = 0;
    const uint32_t *W = words + env.base.n_fse;
    uint32_t cur = 0, literal_index = 0;
    Match pending = {0, 0, 0};
    // Every lane walks its own stream, so a plain load per position costs a memory round trip per step of this serial
    // loop (and prefetch.global.L1 does not shorten it).  Each lane therefore owns a ring of words in shared memory that
    // cp.async keeps filled ahead of its cursor; the cursor reads four words at a time into registers.
    //
    // How the ring knows that a chunk has arrived -- three versions:
    //  1. 16 single-chunk groups in flight per lane, `wait_group 14` before reading ("my 16th newest group is complete"):
    //     right by PTX's per-thread wording, but the lanes are divergent and the hardware keeps ONE group counter per warp;
    //     about ten of 16 384 streams per run read a chunk before it had arrived (valid frames, but not the reference's
    //     bytes, and different ones from run to run; found by scripts/enc_words_diff.py).  11.1 ms per GiB, wrong.
    //  2. two half-ring bursts per lane, `wait_group 0` only: correct, 14.3 ms -- half of all warp instructions were the
    //     request loops, executed by 1.9 lanes at a time, and every drain also waited for the other lanes' fresh requests.
    //  3. (this one) the upkeep runs at the top of the loop for the whole, converged warp: every lane requests at most two
    //     chunks, ONE group is committed per iteration, and `wait_group kLag - 1` then means exactly "everything requested
    //     kLag iterations ago has arrived" -- per warp, which is how the hardware counts.  Every kLag iterations a lane moves
    //     its `safe` mark to what it had requested kLag..2 kLag iterations earlier; a lane whose cursor gets ahead of its
    //     mark (a long jump) asks for a drain, which the whole warp then executes.
    __shared__ __align__(16) uint8_t rings[kReplayThreads * kRingStride];
    constexpr uint32_t kLag = 8;
    static_assert((kRingWords & (kRingWords - 1)) == 0 && (kLag & (kLag - 1)) == 0, "powers of two");
    const uint32_t ring = (uint32_t)__cvta_generic_to_shared(rings) + threadIdx.x * kRingStride;
    const uint32_t w_limit = valid ? (end + 3u) & ~3u : 0u;  // chunks at or beyond this word index are never needed
    uint32_t wbase = 0xFFFFFFFFu, fetched = 0, safe = 0, mark = 0, iter = 0;
    uint4 wq = make_uint4(0, 0, 0, 0);
    auto request = [&](bool go) {  // one chunk, predicated: the instruction is the same for all 32 lanes
        const bool p = go && fetched < w_limit;
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p cp.async.cg.shared.global [%0], [%1], 16;\n\t}"
                     ::"r"(ring + (fetched & (kRingWords - 1)) * 4), "l"(W + (p ? fetched : 0u)), "r"((uint32_t)p) : "memory");
        fetched += p ? 4u : 0u;
    };
    for (;;) {
        __syncwarp();
        active = active && cur < end;
        if (!__any_sync(0xFFFFFFFFu, active)) break;
        {   // ---- ring upkeep, whole warp ----
            const uint32_t wb = cur & ~3u;
            // a cursor that jumped past everything requested restarts the ring there; what is still in flight must land first
            const bool restart = active && wb > fetched;
            bool drain = __any_sync(0xFFFFFFFFu, restart);
            if (drain) asm volatile("cp.async.wait_group 0;" ::: "memory");
            if (restart) fetched = wb;
            if (drain) { safe = fetched; mark = fetched; }   // (for a restarted lane: nothing of the new range yet)
            const bool room = active && fetched < wb + kRingWords;
            request(room);
            request(active && fetched < wb + kRingWords);
            if (iter == 0) {  // a few more at the start, while the warp is converged anyway
#pragma unroll
                for (int k = 0; k < 6; k++) request(active && fetched < wb + kRingWords);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group %0;" ::"n"(kLag - 1) : "memory");
            if ((iter & (kLag - 1)) == 0) { safe = mark; mark = fetched; }  // requested >= kLag iterations ago: arrived
            const bool need = active && wb + 4 > safe;
            if (__any_sync(0xFFFFFFFFu, need)) {
                asm volatile("cp.async.wait_group 0;" ::: "memory");
                safe = fetched; mark = fetched;
            }
            iter++;
        }
        if (!active) continue;
        if ((cur & ~3u) != wbase) {
            wbase = cur & ~3u;
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(wq.x), "=r"(wq.y), "=r"(wq.z), "=r"(wq.w) : "r"(ring + (wbase & (kRingWords - 1)) * 4) : "memory");
        }
        const uint32_t k4 = cur & 3u;
        const uint32_t w = k4 == 0 ? wq.x : (k4 == 1 ? wq.y : (k4 == 2 ? wq.z : wq.w));
        if ((w & 0x3FFFFu) == 0) { cur += (w >> 18) & 0x3FFu; continue; }  // find_match came back empty (:203-209) for this many positions
        Match inc;
        inc.idx = cur;
        inc.match_idx = cur - (w & 0x3FFFFu);
        inc.match_len = (w >> 18) & 0x3FFu;
        if (inc.match_len == kWordLenSat) {  // the word's length field is saturated: finish the extension here
            const uint32_t maxl = len - cur;
            while (inc.match_len + 8 <= maxl) {
                const uint64_t y = ld8u(src + cur + inc.match_len) ^ ld8u(src + inc.match_idx + inc.match_len);
                if (y) { inc.match_len += (__ffsll((long long)y) - 1) >> 3; goto fwd_done; }
                inc.match_len += 8;
            }
            while (inc.match_len < maxl && src[cur + inc.match_len] == src[inc.match_idx + inc.match_len]) inc.match_len++;
        fwd_done:;
        }
        {   // match_dec (:261-268)
            const uint32_t lit = cur - literal_index;
            const uint32_t lim = lit < inc.match_idx ? lit : inc.match_idx;
            const uint32_t bw = w >> 28;
            uint32_t dec = bw < lim ? bw : lim;
            if (bw == kWordBwSat) while (dec < lim && src[inc.idx - dec - 1] == src[inc.match_idx - dec - 1]) dec++;
            inc.idx -= dec; inc.match_idx -= dec; inc.match_len += dec;
        }
        bool have = true;
        Match sel = pending;
        if (inc.match_len >= kGoodMatchLen) { sel = inc; pending.match_len = 0; }
        else if (pending.match_len == 0) { pending = inc; have = false; }
        else if ((int32_t)(pending.idx + pending.match_len - inc.idx) <= 0) { pending = inc; }
        else if (inc.match_len > pending.match_len) { sel = inc; pending.match_len = 0; }
        else { pending.match_len = 0; }
        if (have) {
            tsink_push_match(fs, env, sel.idx - literal_index, sel.match_len, sel.idx - sel.match_idx);
            literal_index = sel.idx + sel.match_len;
            if (literal_index >= end) { cur = end; continue; }  // done (the lane stays in the loop: the ring upkeep is warp-wide)
            cur = cur + 1 > literal_index ? cur + 1 : literal_index;
        } else {
            cur++;
        }
    }
    if (!valid) return;
    if (pending.match_len != 0) {
        tsink_push_match(fs, env, pending.idx - literal_index, pending.match_len, pending.idx - pending.match_idx);
        literal_index = pending.idx + pending.match_len;
    }
    if (len - literal_index != 0) tsink_push_match(fs, env, len - literal_index, 0, 1);
    tsink_emit_block(fs, env);
    streams[si].n_blocks = fs.n_blocks;
}

#include "encode_long.cuh"

// ------------------------------------------------------------------------------------------------
// FSE block encode: warp per block.
//
// The seven FSE state machines of a block (4 literal states, L, M, D) are independent chains: each
// only sees its own symbol sequence (fse/encoder.rs:190-200).  What couples them is the position of
// their bits in the stream, and that is a prefix sum.  So per chunk (128 literals / 32 packs) a few
// lanes run the chains and record (bits, count) per symbol, then all 32 lanes concatenate in the
// reference's emission order (fse/literals.rs:109-121, fse/lmds.rs:74-87) with a warp scan and OR the
// bits into a shared-memory bit buffer that is flushed to the block as whole bytes.
// ------------------------------------------------------------------------------------------------
constexpr int kFseEncWarps = 8;
struct FseEncSmem {
    uint32_t W[360];       // histogram, then normalised weights
    uint32_t E[360];       // encode table: t_k (low 16) | t_w (high 16)
    uint32_t bitbuf[72];   // chunk bit buffer (<= 31 carried + 32 * 54 bits)
    uint16_t chain[128];   // (count | bits << 4) per symbol, in emission order
    int2 ent[128];         // the chunk's table entries {t_k, t_w}, looked up and unpacked by all lanes before the chain lanes run
};

__device__ __forceinline__ uint32_t l_sym(uint32_t v) { return v < 16 ? v : (v < 20 ? 16u : (v < 28 ? 17u : (v < 60 ? 18u : 19u))); }   // L_BASE_FROM_VALUE
__device__ __forceinline__ uint32_t m_sym(uint32_t v) { return v < 16 ? v : (v < 24 ? 16u : (v < 56 ? 17u : (v < 312 ? 18u : 19u))); }  // M_BASE_FROM_VALUE
__device__ __forceinline__ uint32_t d_sym(uint32_t v) {  // largest symbol with D_BASE_VALUE <= v (d_index + D_BASE_FROM_VALUE)
    if (v < 4) return v;
    const uint32_t e = 29 - __clz(v + 4);  // base(4e) = 2^(e+2) - 4 <= v
    return 4 * e + (((v + 4) >> e) - 4);
}
__device__ __forceinline__ uint32_t d_base_e(uint32_t s) { return ((4u + (s & 3u)) << (s >> 2)) - 4u; }
__device__ __forceinline__ uint32_t l_extra_e(uint32_t s) { return s < 16 ? 0u : (s == 16 ? 2u : (s == 17 ? 3u : (s == 18 ? 5u : 8u))); }
__device__ __forceinline__ uint32_t m_extra_e(uint32_t s) { return s < 16 ? 0u : (s == 16 ? 3u : (s == 17 ? 5u : (s == 18 ? 8u : 11u))); }
__device__ __forceinline__ uint32_t l_base_e(uint32_t s) { return s < 16 ? s : ((0x3C1C1410u >> ((s - 16) * 8)) & 0xFFu); }
__device__ __forceinline__ uint32_t m_base_e(uint32_t s) { return s < 16 ? s : (uint32_t)((0x0138003800180010ull >> ((s - 16) * 16)) & 0xFFFFu); }

// normalize_m1 (fse/weights.rs:218-278), one thread.
__device__ void normalize_m1(uint32_t *w, uint32_t n_sym, uint32_t in_total, uint32_t out_total) {
    int32_t remaining = 0;
    uint32_t max_index = 0;
    if (in_total != 0) {
        const uint32_t shift = __clz(out_total), multiply = (1u << 31) / in_total, round = 1u << (shift - 1);
        uint32_t max_weight = 0;
        remaining = (int32_t)out_total;
        for (uint32_t i = 0; i < n_sym; i++) {
            const uint32_t v = w[i];
            if (v == 0) continue;
            uint32_t f = (v * multiply + round) >> shift;
            if (f == 0) f = 1;
            w[i] = f;
            remaining -= (int32_t)f;
            if (f > max_weight) { max_weight = f; max_index = i; }
        }
    }
    if (-remaining < (int32_t)w[max_index] / 4) {
        w[max_index] = (uint32_t)((int32_t)w[max_index] + remaining);
    } else {
        uint32_t overflow = (uint32_t)(-remaining);
        for (int shift = 3; shift >= 0; shift--)
            for (uint32_t i = 0; i < n_sym; i++) {
                if (overflow == 0) break;
                const uint32_t v = w[i];
                if (v == 0) continue;
                uint32_t k = (v - 1) >> shift;
                if (k > overflow) k = overflow;
                w[i] = v - k;
                overflow -= k;
            }
    }
}
// build_e_table (fse/encoder.rs:219-240), one thread
__device__ void build_e_table(const uint32_t *w, uint32_t *e, uint32_t n_sym, uint32_t n_states) {
    const uint32_t n_clz = __clz(n_states);
    uint32_t total = 0;
    for (uint32_t i = 0; i < n_sym; i++) {
        const uint32_t v = w[i];
        int32_t t_k, t_w;
        if (v == 0) { t_k = -(int32_t)n_states; t_w = 0; }
        else {
            const uint32_t k = __clz(v) - n_clz;
            t_k = (int32_t)(1024 * k) - (int32_t)(v << k);
            t_w = (int32_t)n_states + (int32_t)total - (int32_t)v;
        }
        e[i] = ((uint32_t)t_k & 0xFFFFu) | ((uint32_t)t_w << 16);
        total += v;
    }
}
// EEntry::encode (fse/encoder.rs:190-200): returns count | bits << 4
__device__ __forceinline__ int2 e_unpack(uint32_t entry) { return make_int2((int16_t)(entry & 0xFFFF), (int16_t)(entry >> 16)); }
__device__ __forceinline__ uint32_t e_step2(int2 e, uint32_t &state) {  // the same with the entry already unpacked
    const uint32_t s = state;
    const uint32_t nb = (uint32_t)(e.x + (int32_t)s) >> 10;
    state = (uint32_t)(e.y + (int32_t)(s >> nb));
    return nb | ((s & ((1u << nb) - 1u)) << 4);
}
__device__ __forceinline__ uint32_t e_step(uint32_t entry, uint32_t &state) {
    const int32_t t_k = (int16_t)(entry & 0xFFFF), t_w = (int16_t)(entry >> 16);
    const uint32_t s = state;
    const uint32_t nb = (uint32_t)(t_k + (int32_t)s) >> 10;
    state = (uint32_t)(t_w + (int32_t)(s >> nb));
    return nb | ((s & ((1u << nb) - 1u)) << 4);
}
// Warp-wide append of one value per lane (lane order = stream order) to the block's byte stream.
// `carry` (< 8 bits, value in carry_val) is what the previous chunk left over.
__device__ __forceinline__ void warp_emit(uint32_t *bitbuf, uint64_t v, uint32_t n, uint8_t *&out, uint32_t &carry, uint32_t &carry_val, uint32_t lane) {
    for (uint32_t t = lane; t < 72; t += 32) bitbuf[t] = t == 0 ? carry_val : 0u;
    uint32_t inc = n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t x = __shfl_up_sync(0xFFFFFFFFu, inc, o);
        if (lane >= (uint32_t)o) inc += x;
    }
    const uint32_t total = carry + __shfl_sync(0xFFFFFFFFu, inc, 31);
    __syncwarp();
    if (n) {
        const uint32_t pos = carry + inc - n, w = pos >> 5, sh = pos & 31;
        const uint64_t lo = v << sh;
        atomicOr(&bitbuf[w], (uint32_t)lo);
        if (sh + n > 32) atomicOr(&bitbuf[w + 1], (uint32_t)(lo >> 32));
        if (sh + n > 64) atomicOr(&bitbuf[w + 2], (uint32_t)(v >> (64 - sh)));
    }
    __syncwarp();
    const uint32_t n_bytes = total >> 3;
    const uint8_t *bb = reinterpret_cast<const uint8_t *>(bitbuf);
    for (uint32_t t = lane; t < n_bytes; t += 32) out[t] = bb[t];
    out += n_bytes;
    carry = total & 7;
    carry_val = carry ? (uint32_t)bb[n_bytes] & ((1u << carry) - 1u) : 0u;
    __syncwarp();
}

__global__ void __launch_bounds__(kFseEncWarps * 32)
k_enc_fse_blocks(EncBlock *blocks, const uint32_t *__restrict__ n_blocks_p, const uint2 *__restrict__ pack_scratch,
                 uint8_t *lit_scratch, uint8_t *out_scratch, uint32_t *work_counter, const uint8_t *__restrict__ src_base) {
    __shared__ FseEncSmem sm_all[kFseEncWarps];
    FseEncSmem &sm = sm_all[threadIdx.x >> 5];
    const uint32_t lane = lane_id();
    const uint32_t n_blocks = *n_blocks_p;
    for (;;) {
        uint32_t bi = 0;
        if (lane == 0) bi = atomicAdd(work_counter, 1u);
        bi = __shfl_sync(0xFFFFFFFFu, bi, 0);
        if (bi >= n_blocks) break;
        const EncBlock b = blocks[bi];
        const uint2 *packs = pack_scratch + b.pack_off;
        uint8_t *lits = lit_scratch + b.lit_off;
        uint8_t *out = out_scratch + b.out_off;
        if (b.gather) {
            // Blocks of k_enc_replay: the literal bytes are still in the source.  Pack i's literals start at the block's
            // first source byte + the (L + M) of the packs before it (Buffer::push appends them in that order).
            const uint8_t *sp = src_base + b.src_pos;
            uint32_t lit_run = 0, src_run = 0;
            for (uint32_t i0 = 0; i0 < b.n_packs; i0 += 32) {
                uint32_t L = 0, M = 0;
                if (i0 + lane < b.n_packs) { const uint2 p = packs[i0 + lane]; L = p.x & 0xFFFF; M = p.x >> 16; }
                const uint32_t v = (L << 17) + (L + M);  // 32 * 315 < 2^14, 32 * (315 + 2359) < 2^17
                uint32_t inc = v;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
                    if (lane >= (uint32_t)o) inc += t;
                }
                const uint32_t exc = inc - v, tot = __shfl_sync(0xFFFFFFFFu, inc, 31);
                const uint32_t my_lit = lit_run + (exc >> 17), my_src = src_run + (exc & 0x1FFFF);
                const uint32_t sl = L <= 16 ? L : 0;
                const uint32_t max_l = __reduce_max_sync(0xFFFFFFFFu, sl);
                for (uint32_t t = 0; t < max_l; t++)
                    if (t < sl) lits[my_lit + t] = sp[my_src + t];
                uint32_t mask = __ballot_sync(0xFFFFFFFFu, L > 16);
                while (mask) {
                    const int j = __ffs(mask) - 1;
                    mask &= mask - 1;
                    const uint32_t o = __shfl_sync(0xFFFFFFFFu, my_lit, j), f = __shfl_sync(0xFFFFFFFFu, my_src, j), nn = __shfl_sync(0xFFFFFFFFu, L, j);
                    for (uint32_t t = lane; t < nn; t += 32) lits[o + t] = sp[f + t];
                }
                lit_run += tot >> 17; src_run += tot & 0x1FFFF;
            }
            __syncwarp();
        }
        // ---- Weights::load (fse/weights.rs:25-64): histograms of the real packs / literals ----
        for (uint32_t t = lane; t < 360; t += 32) sm.W[t] = 0;
        __syncwarp();
        for (uint32_t i = lane; i < b.n_packs; i += 32) {
            const uint2 p = packs[i];
            atomicAdd(&sm.W[l_sym(p.x & 0xFFFF)], 1u);
            atomicAdd(&sm.W[20 + m_sym(p.x >> 16)], 1u);
            atomicAdd(&sm.W[40 + d_sym(p.y)], 1u);
        }
        for (uint32_t i = lane; i < b.n_lits; i += 32) atomicAdd(&sm.W[104 + lits[i]], 1u);
        __syncwarp();
        // ---- normalize_m1 x4, one table per lane ----
        if (lane == 0 && b.n_packs) normalize_m1(sm.W, 20, b.n_packs, kLStates);
        if (lane == 1 && b.n_packs) normalize_m1(sm.W + 20, 20, b.n_packs, kMStates);
        if (lane == 2 && b.n_packs) normalize_m1(sm.W + 40, 64, b.n_packs, kDStates);
        if (lane == 3 && b.n_lits) normalize_m1(sm.W + 104, 256, b.n_lits, kUStates);
        __syncwarp();
        // ---- Encoder::init: one table per lane ----
        if (lane == 0) build_e_table(sm.W, sm.E, 20, kLStates);
        if (lane == 1) build_e_table(sm.W + 20, sm.E + 20, 20, kMStates);
        if (lane == 2) build_e_table(sm.W + 40, sm.E + 40, 64, kDStates);
        if (lane == 3) build_e_table(sm.W + 104, sm.E + 104, 256, kUStates);
        // ---- Weights::store_v2 (fse/weights.rs:139-163, fse/weight_encoder.rs:23-37), 32 weights per step ----
        uint8_t *wp = out + kV2HeaderSize;
        uint32_t carry = 0, carry_val = 0;
        for (uint32_t base = 0; base < 360; base += 32) {
            const uint32_t i = base + lane;
            uint32_t u = 0, ub = 0;
            if (i < 360) {
                const uint32_t v = sm.W[i];
                if (v == 0) { u = 0; ub = 2; } else if (v == 1) { u = 2; ub = 2; } else if (v == 2) { u = 1; ub = 3; } else if (v == 3) { u = 5; ub = 3; }
                else if (v < 8) { u = 3 + ((v - 4) << 3); ub = 5; } else if (v < 24) { u = ((v - 8) << 4) + 7; ub = 8; } else { u = ((v - 24) << 4) + 15; ub = 14; }
            }
            warp_emit(sm.bitbuf, u, ub, wp, carry, carry_val, lane);
        }
        if (carry) { if (lane == 0) *wp = (uint8_t)carry_val; wp++; }
        const uint32_t n_weight_bytes = (uint32_t)(wp - (out + kV2HeaderSize));
        __syncwarp();
        // ---- Literals::store (fse/literals.rs:93-133): padded to x4 with literals[0], last to first ----
        const uint32_t n_lit_pad = (b.n_lits + 3) / 4 * 4;
        uint8_t *lp = wp;
        carry = 0; carry_val = 0;
        uint32_t ust = kUStates;  // lanes 0..3 hold literal states 0..3
        const uint32_t pad = b.n_lits ? lits[0] : 0;
        for (uint32_t hi = n_lit_pad; hi != 0;) {
            const uint32_t n = hi < 128 ? hi : 128;  // literals [hi - n, hi), emitted from hi - 1 downwards
            for (uint32_t t = lane; t < n; t += 32) {  // emission slot t holds literal hi - 1 - t
                const uint32_t idx = hi - 1 - t;
                sm.ent[t] = e_unpack(sm.E[104 + (idx < b.n_lits ? lits[idx] : pad)]);
            }
            __syncwarp();
            if (lane < 4) {  // literal idx uses state idx & 3 (the loop encodes i-1 with state 3 ... i-4 with state 0)
                const uint32_t first = 3 - lane;  // slots first, first + 4, ... belong to this state
#pragma unroll 4
                for (uint32_t t = first; t < n; t += 4) sm.chain[t] = (uint16_t)e_step2(sm.ent[t], ust);
            }
            __syncwarp();
            {   // each lane packs four consecutive slots
                uint64_t v = 0; uint32_t nb = 0;
#pragma unroll
                for (uint32_t k = 0; k < 4; k++) {
                    const uint32_t t = lane * 4 + k;
                    if (t < n) { const uint32_t c = sm.chain[t]; v |= (uint64_t)(c >> 4) << nb; nb += c & 15; }
                }
                warp_emit(sm.bitbuf, v, nb, lp, carry, carry_val, lane);
            }
            hi -= n;
        }
        uint32_t lit_bits = 0;
        if (carry) { if (lane == 0) *lp = (uint8_t)carry_val; lp++; lit_bits = 8 - carry; }
        const uint32_t n_lit_payload = (uint32_t)(lp - wp);
        const uint32_t us0 = __shfl_sync(0xFFFFFFFFu, ust, 0), us1 = __shfl_sync(0xFFFFFFFFu, ust, 1), us2 = __shfl_sync(0xFFFFFFFFu, ust, 2),
                       us3 = __shfl_sync(0xFFFFFFFFu, ust, 3);
        // ---- Lmds::store (fse/lmds.rs:62-93): 8 zero bytes, then D, M, L of each pack, last pack first ----
        uint8_t *mp = lp;
        if (lane < 8) mp[lane] = 0;
        mp += 8;
        carry = 0; carry_val = 0;
        uint32_t st = lane == 0 ? kLStates : (lane == 1 ? kMStates : kDStates);  // lanes 0,1,2 hold the L, M, D states
        for (uint32_t hi = b.n_packs; hi != 0;) {
            const uint32_t n = hi < 32 ? hi : 32;  // packs [hi - n, hi); slot t holds pack hi - 1 - t
            uint32_t l = 0, m = 0, d = 0, sl = 0, smm = 0, sd = 0;
            if (lane < n) {
                const uint2 p = packs[hi - 1 - lane];
                l = p.x & 0xFFFF; m = p.x >> 16; d = p.y;
                sl = l_sym(l); smm = m_sym(m); sd = d_sym(d);
                sm.ent[lane] = e_unpack(sm.E[sl]); sm.ent[32 + lane] = e_unpack(sm.E[20 + smm]); sm.ent[64 + lane] = e_unpack(sm.E[40 + sd]);
            }
            __syncwarp();
            if (lane < 3) {
#pragma unroll 4
                for (uint32_t t = 0; t < n; t++) sm.chain[lane * 32 + t] = (uint16_t)e_step2(sm.ent[lane * 32 + t], st);
            }
            __syncwarp();
            uint64_t v = 0; uint32_t nb = 0;
            if (lane < n) {  // D extra, D state, M extra, M state, L extra, L state
                const uint32_t cl = sm.chain[lane], cm = sm.chain[32 + lane], cd = sm.chain[64 + lane];
                v = d - d_base_e(sd); nb = sd >> 2;
                v |= (uint64_t)(cd >> 4) << nb; nb += cd & 15;
                v |= (uint64_t)(m - m_base_e(smm)) << nb; nb += m_extra_e(smm);
                v |= (uint64_t)(cm >> 4) << nb; nb += cm & 15;
                v |= (uint64_t)(l - l_base_e(sl)) << nb; nb += l_extra_e(sl);
                v |= (uint64_t)(cl >> 4) << nb; nb += cl & 15;
            }
            warp_emit(sm.bitbuf, v, nb, mp, carry, carry_val, lane);
            hi -= n;
        }
        uint32_t lmd_bits = 0;
        if (carry) { if (lane == 0) *mp = (uint8_t)carry_val; mp++; lmd_bits = 8 - carry; }
        const uint32_t n_lmd_payload = (uint32_t)(mp - lp);
        const uint32_t fl = __shfl_sync(0xFFFFFFFFu, st, 0), fm = __shfl_sync(0xFFFFFFFFu, st, 1), fd = __shfl_sync(0xFFFFFFFFu, st, 2);
        // ---- FseBlock::store_v2 (fse/block.rs:168-196) ----
        if (lane == 0) {
            const uint32_t n_raw = b.n_lits + b.n_match_bytes;
            uint64_t h[4];
            h[0] = (uint64_t)kMagicVx2 | ((uint64_t)n_raw << 32);
            h[1] = (uint64_t)n_lit_pad | ((uint64_t)n_lit_payload << 20) | ((uint64_t)b.n_packs << 40) | ((uint64_t)(7 - lit_bits) << 60);
            h[2] = (uint64_t)(us0 - kUStates) | ((uint64_t)(us1 - kUStates) << 10) | ((uint64_t)(us2 - kUStates) << 20) | ((uint64_t)(us3 - kUStates) << 30) |
                   ((uint64_t)n_lmd_payload << 40) | ((uint64_t)(7 - lmd_bits) << 60);
            h[3] = (uint64_t)(kV2HeaderSize + n_weight_bytes) | ((uint64_t)(fl - kLStates) << 32) | ((uint64_t)(fm - kMStates) << 42) |
                   ((uint64_t)(fd - kDStates) << 52);
            for (int k = 0; k < 32; k++) out[k] = (uint8_t)(h[k >> 3] >> (8 * (k & 7)));
            blocks[bi].out_size = (uint32_t)(mp - out);
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// assemble: warp per stream (frontend_bytes.rs:50-111: block selection, raw fallback, bvx$)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void warp_copy_bytes(uint8_t *dst, const uint8_t *src, uint64_t n, uint32_t lane) {
    const uint64_t head = (16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15;
    if (((reinterpret_cast<uintptr_t>(dst) ^ reinterpret_cast<uintptr_t>(src)) & 15) == 0 && n >= head + 16) {
        if (lane < head) dst[lane] = src[lane];
        const uint64_t nv = (n - head) / 16;
        const uint4 *s4 = reinterpret_cast<const uint4 *>(src + head);
        uint4 *d4 = reinterpret_cast<uint4 *>(dst + head);
        for (uint64_t i = lane; i < nv; i += 32) d4[i] = s4[i];
        for (uint64_t i = head + nv * 16 + lane; i < n; i += 32) dst[i] = src[i];
    } else {
        for (uint64_t i = lane; i < n; i += 32) dst[i] = src[i];
    }
}
__device__ __forceinline__ void put_u32(uint8_t *p, uint32_t v) { for (int k = 0; k < 4; k++) p[k] = (uint8_t)(v >> (8 * k)); }

constexpr int kAsmWarps = 8;
__global__ void __launch_bounds__(kAsmWarps * 32)
k_enc_assemble(const uint8_t *__restrict__ src_base, const uint64_t *__restrict__ src_off, const uint64_t *__restrict__ src_len,
               uint8_t *__restrict__ dst_base, const uint64_t *__restrict__ dst_off, const uint64_t *__restrict__ dst_cap, size_t n_streams,
               const EncStream *__restrict__ streams, const StreamCounts *__restrict__ bases, const uint32_t *__restrict__ block_ids,
               EncBlock *blocks, const uint8_t *__restrict__ out_scratch, uint64_t *out_len, int32_t *status) {
    const uint32_t lane = lane_id();
    const size_t si = (size_t)blockIdx.x * kAsmWarps + (threadIdx.x >> 5);
    if (si >= n_streams) return;
    if (status[si] != LZFSE_B200_OK) { if (lane == 0) out_len[si] = 0; return; }
    const EncStream st = streams[si];
    const uint8_t *src = src_base + src_off[si];
    const uint64_t len = src_len[si], cap = dst_cap[si];
    uint8_t *dst = dst_base + dst_off[si];
    uint64_t total = 4;  // bvx$
    bool raw = st.kind == SK_RAW;
    if (st.kind == SK_VN) {
        // frontend_bytes.rs:92-99: an LZVN block that is not smaller than a raw one is redone as raw
        if (len < kRawLimit && len + 8 <= st.vn_size) raw = true;
        else total += st.vn_size;
    } else if (st.kind == SK_FSE) {
        uint64_t part = 0;
        for (uint32_t k = lane; k < st.n_blocks; k += 32) part += blocks[block_ids[bases[si].n_literals + k]].out_size;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xFFFFFFFFu, part, o);
        total += part;
    }
    if (raw) total += 8 + len;
    if (total > cap) {  // the reference would grow its Vec; a fixed buffer reports BufferOverflow
        if (lane == 0) { status[si] = LZFSE_B200_BUFFER_OVERFLOW; out_len[si] = 0; }
        return;
    }
    uint64_t pos = 0;
    if (raw) {
        if (lane == 0) { put_u32(dst, kMagicRaw); put_u32(dst + 4, (uint32_t)len); }
        warp_copy_bytes(dst + 8, src, len, lane);
        pos = 8 + len;
    } else if (st.kind == SK_VN) {
        warp_copy_bytes(dst, out_scratch + bases[si].n_lmds, st.vn_size, lane);
        pos = st.vn_size;
    } else if (st.fast == 2) {
        // a long stream's blocks are copied side by side by k_long_copy; here only where each of them goes
        for (uint32_t k0 = 0; k0 < st.n_blocks; k0 += 32) {
            const uint32_t k = k0 + lane;
            const uint32_t id = k < st.n_blocks ? block_ids[bases[si].n_literals + k] : 0u;
            const uint32_t sz = k < st.n_blocks ? blocks[id].out_size : 0u;
            uint64_t inc = sz;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint64_t t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
                if (lane >= (uint32_t)o) inc += t;
            }
            if (k < st.n_blocks) { blocks[id].dst_pos = dst_off[si] + pos + inc - sz; blocks[id].pad = 1; }
            pos += __shfl_sync(0xFFFFFFFFu, inc, 31);
        }
    } else {
        for (uint32_t k = 0; k < st.n_blocks; k++) {
            const EncBlock b = blocks[block_ids[bases[si].n_literals + k]];
            warp_copy_bytes(dst + pos, out_scratch + b.out_off, b.out_size, lane);
            pos += b.out_size;
        }
    }
    if (lane == 0) { put_u32(dst + pos, kMagicEos); out_len[si] = total; }
}

}  // namespace lzb

using namespace lzb;

struct lzfse_b200_encoder {
    int device = 0;
    int n_sms = 148;
    cudaStream_t own_stream = nullptr;
    std::string last_error;
    uint64_t launches = 0;
    DevBuf streams, counts, totals_dev, tables, packs, lits, block_ids, blocks, out, counters, words;
    DevBuf long_list, seg_list, l_prev, l_heads, l_cseg, l_rseg, l_out, l_spec, l_fix, l_states, l_tail, l_agg, l_csum, l_entry;  // long streams (encode_long.cuh)
    int allow_long = 1;  // LZB_ENC_LONG=0 sends streams > 64 KiB through k_enc_parse (measurements, tests)
    int allow_seg = 1;   // LZB_ENC_SEG=0: streams <= 64 KiB are replayed whole by k_enc_replay (one thread per stream) instead of per segment
    uint32_t last_long_redo = 0;
    bool pending = false;             // an *_async call has been enqueued and not yet synchronised
    cudaStream_t pending_stream = nullptr;
    bool tables_dirty = false;        // a call that ran k_enc_parse did not complete: entries of unknown epochs may be in the tables
    int allow_fast = 1;  // LZB_ENC_FAST=0 sends every stream through k_enc_parse (measurements, tests)
    PinnedBuf totals_host;
    HostStage stage;
    StageTimer timer;
};

namespace {

constexpr int kParseWarpsPerSm = LZB_PARSE_CTAS * 4;  // resident history tables: 148 * 28 * 512 KiB = 2072 MiB

int encode_batch_device_impl(lzfse_b200_encoder *e, const uint8_t *src, const uint64_t *src_off, const uint64_t *src_len, uint8_t *dst,
                             const uint64_t *dst_off, const uint64_t *dst_cap, uint64_t *out_len, int32_t *status, size_t n, cudaStream_t s,
                             bool wait = true) {
    e->launches = 0;
    e->pending = false;
    if (n == 0) return LZFSE_B200_OK;
    if (n > 0x7FFFFFFFull) { e->last_error = "too many streams in one batch"; return LZFSE_B200_INVALID_ARGUMENT; }
    LZB_CK(e, e->streams.reserve(n * sizeof(EncStream)));
    LZB_CK(e, e->counts.reserve((n + 1 + n / 1024 + 2) * sizeof(StreamCounts)));  // + tile sums of the scan
    LZB_CK(e, e->totals_dev.reserve(sizeof(StreamCounts)));
    LZB_CK(e, e->totals_host.reserve(2 * sizeof(StreamCounts)));
    LZB_CK(e, e->counters.reserve(16 * sizeof(uint32_t)));
    LZB_CK(e, e->long_list.reserve(n * sizeof(uint32_t)));
    LZB_CK(e, e->seg_list.reserve(n * sizeof(uint32_t)));
    const int tb = 128;
    e->timer.begin(s);
    LZB_CK(e, cudaMemsetAsync(e->counters.p, 0, 16 * sizeof(uint32_t), s));
    // [0] blocks produced, [1] parse cursor, [2] fse-encode cursor, [3] find cursor, [4] fast streams, [5] k_enc_parse streams,
    // [6] long streams, [7] replay segments, [8] chain pieces, [10..11] prev[] elements, [12] chain cursor, [13] segments stitched again, [14] segments whose packs were pushed one by one
    uint32_t *ctr = e->counters.as<uint32_t>();
    k_enc_prep<<<(unsigned)((n + tb - 1) / tb), tb, 0, s>>>(src_len, n, e->streams.as<EncStream>(), e->counts.as<StreamCounts>(), status, ctr + 4,
                                                           e->long_list.as<uint32_t>(), e->seg_list.as<uint32_t>(), e->allow_fast, e->allow_long, e->allow_seg);
    launch_exclusive_scan(e->counts.as<StreamCounts>(), n, e->totals_host.as<StreamCounts>(), nullptr, s);  // pinned host memory (UVA)
    k_enc_publish_counts<<<1, 1, 0, s>>>(ctr + 4, reinterpret_cast<uint32_t *>(e->totals_host.as<StreamCounts>() + 1));
    e->launches += n > 8192 ? 5 : 3;  // prep + the exclusive scan (three launches for large batches) + publish
    LZB_CK(e, cudaStreamSynchronize(s));
    const StreamCounts tot = *e->totals_host.as<StreamCounts>();  // {packs, literal bytes, block slots, out bytes}
    const uint32_t n_fast = reinterpret_cast<const uint32_t *>(e->totals_host.as<StreamCounts>() + 1)[0];
    const uint32_t n_slow = reinterpret_cast<const uint32_t *>(e->totals_host.as<StreamCounts>() + 1)[1];
    const uint32_t n_long = reinterpret_cast<const uint32_t *>(e->totals_host.as<StreamCounts>() + 1)[2];
    const uint32_t n_rseg = reinterpret_cast<const uint32_t *>(e->totals_host.as<StreamCounts>() + 1)[3];
    const uint32_t n_cseg = reinterpret_cast<const uint32_t *>(e->totals_host.as<StreamCounts>() + 1)[4];
    const uint32_t n_segl = reinterpret_cast<const uint32_t *>(e->totals_host.as<StreamCounts>() + 1)[5];
    const uint64_t long_elems = reinterpret_cast<const uint64_t *>(e->totals_host.as<StreamCounts>() + 1)[3];
    if (tot.n_literals > 0xFFFFFFF0ull) { e->last_error = "too many blocks in one batch"; return LZFSE_B200_INVALID_ARGUMENT; }
    const unsigned parse_ctas = (unsigned)e->n_sms * kParseWarpsPerSm / kParseWarps;
    const size_t n_slots = (size_t)parse_ctas * kParseWarps;
    if (n_slow) {   // history tables + one epoch word per table; zeroed when (re)allocated, never again (see k_enc_parse)
        const void *before = e->tables.p;
        LZB_CK(e, e->tables.reserve(n_slots * (kTableWords + 1) * sizeof(uint32_t)));
        // The epochs are written back when a stream's parse ends; after a call that did not get that far (a CUDA error, a
        // fault) the tables may hold positions biased with epochs the epoch words do not know about, which a later stream
        // could take for valid candidates.  Such a call leaves the flag set and the next one starts from zeroed tables.
        if (e->tables.p != before || e->tables_dirty) LZB_CK(e, cudaMemsetAsync(e->tables.p, 0, e->tables.cap, s));
        e->tables_dirty = true;
    }
    if (n_fast || n_long) LZB_CK(e, e->words.reserve((tot.n_fse + 256) * sizeof(uint32_t)));  // one word per position (k_enc_find -> k_enc_replay)
    if (n_long) {
        LZB_CK(e, e->l_prev.reserve((long_elems + 64) * sizeof(uint2)));   // {chain link, four bytes} per position
        LZB_CK(e, e->l_heads.reserve((size_t)n_cseg * (1u << kHashBits) * sizeof(uint32_t)));
        LZB_CK(e, e->l_cseg.reserve((size_t)n_cseg * sizeof(LongSeg)));
    }
    if (n_segl) {
        LZB_CK(e, e->l_rseg.reserve((size_t)n_rseg * sizeof(LongSeg)));
        LZB_CK(e, e->l_out.reserve((size_t)n_rseg * sizeof(LongSegOut)));
        LZB_CK(e, e->l_spec.reserve((size_t)n_rseg * kEmitCap * sizeof(uint4)));
        LZB_CK(e, e->l_fix.reserve((size_t)n_rseg * kEmitCap * sizeof(uint4)));
        LZB_CK(e, e->l_states.reserve((size_t)n_rseg * kSpecStates * sizeof(uint4)));
        LZB_CK(e, e->l_tail.reserve((size_t)n_segl * 2 * sizeof(uint4)));
        LZB_CK(e, e->l_agg.reserve((size_t)n_rseg * sizeof(SegAgg)));
        LZB_CK(e, e->l_entry.reserve((size_t)n_rseg * sizeof(SegEntry)));
        LZB_CK(e, e->l_csum.reserve((size_t)n_rseg * kSegChunks * sizeof(uint2)));
    }
    LZB_CK(e, e->packs.reserve((tot.n_blocks + 1) * sizeof(uint2)));
    LZB_CK(e, e->lits.reserve(tot.n_fse + 64));
    LZB_CK(e, e->block_ids.reserve((tot.n_literals + 1) * sizeof(uint32_t)));
    LZB_CK(e, e->blocks.reserve((tot.n_literals + 1) * sizeof(EncBlock)));
    LZB_CK(e, e->out.reserve(tot.n_lmds + 64));
    e->timer.mark(s);  // prep

    if (n_fast) {
        const unsigned g = n_fast < (unsigned)e->n_sms ? n_fast : (unsigned)e->n_sms;
        k_enc_find<<<g, kFindThreads, kFindSmemBytes, s>>>(src, src_off, src_len, n, e->streams.as<EncStream>(), e->counts.as<StreamCounts>(),
                                                           e->words.as<uint32_t>(), ctr + 3);
        e->launches += 1;
    }
    e->timer.mark(s);  // find
    if (n_fast && !e->allow_seg) {
        k_enc_replay<<<(unsigned)((n + kReplayThreads - 1) / kReplayThreads), kReplayThreads, 0, s>>>(
            src, src_off, src_len, n, e->streams.as<EncStream>(), e->counts.as<StreamCounts>(), e->words.as<uint32_t>(), e->packs.as<uint2>(),
            e->block_ids.as<uint32_t>(), e->blocks.as<EncBlock>(), ctr);
        e->launches += 1;
    }
    e->timer.mark(s);  // replay
    if (n_slow) {
        k_enc_parse<<<parse_ctas, kParseWarps * 32, 0, s>>>(src, src_off, src_len, n, e->streams.as<EncStream>(), e->counts.as<StreamCounts>(),
                                                           e->tables.as<uint32_t>(), e->tables.as<uint32_t>() + n_slots * kTableWords, e->packs.as<uint2>(),
                                                           e->lits.as<uint8_t>(),
                                                           e->block_ids.as<uint32_t>(), e->blocks.as<EncBlock>(), ctr, e->out.as<uint8_t>(), ctr + 1);
        e->launches += 1;
    }
    e->timer.mark(s);  // parse
    {
        const EncStream *st = e->streams.as<EncStream>();
        const StreamCounts *bs = e->counts.as<StreamCounts>();
        const uint32_t *ll = e->long_list.as<uint32_t>(), *sl = e->seg_list.as<uint32_t>();
        LongSeg *cseg = e->l_cseg.as<LongSeg>(), *rseg = e->l_rseg.as<LongSeg>();
        LongSegOut *so = e->l_out.as<LongSegOut>();
        uint32_t *words = e->words.as<uint32_t>();
        static const bool dbg = getenv("LZB_ENC_DEBUG_SYNC") != nullptr;  // measurements / fault finding: name the kernel that failed
#define LZB_DBG(name) do { if (dbg) { cudaError_t de = cudaStreamSynchronize(s); if (de != cudaSuccess) { e->last_error = std::string(name) + ": " + cudaGetErrorString(de); return LZFSE_B200_CUDA_ERROR; } } } while (0)
        LZB_DBG("before the segment kernels");
        if (n_segl) { k_long_segs<<<n_segl, 64, 0, s>>>(sl, n_segl, st, cseg, rseg); e->launches += 1; }
        LZB_DBG("k_long_segs");
        if (n_long) {
            k_long_heads<<<n_cseg, 512, kHeadsSmem, s>>>(src, src_off, src_len, st, cseg, e->l_heads.as<uint32_t>());
            k_long_carry<<<n_long * ((1u << kHashBits) / 256), 256, 0, s>>>(ll, n_long, st, e->l_heads.as<uint32_t>());
            const unsigned cg = (unsigned)e->n_sms * kChainPerSm;
            k_long_chain<<<n_cseg < cg ? n_cseg : cg, 32, kChainSmem, s>>>(src, src_off, src_len, st, cseg, n_cseg, e->l_prev.as<uint2>(),
                                                                             e->l_heads.as<uint32_t>(), ctr + 12);
            e->launches += 3;
        }
        LZB_DBG("k_long_heads / carry / chain");
        e->timer.mark(s);  // long_chain
        if (n_long) {
            k_long_find<<<n_cseg * (kCSeg / kLFindThreads), kLFindThreads, 0, s>>>(src, src_off, src_len, st, bs, cseg, e->l_prev.as<uint2>(), words);
            e->launches += 1;
        }
        LZB_DBG("k_long_find");
        e->timer.mark(s);  // long_find
        if (n_segl)
            k_long_replay<<<(n_rseg + kReplayThreads - 1) / kReplayThreads, kReplayThreads, 0, s>>>(src, src_off, src_len, bs, rseg, n_rseg, words,
                                                                                                     e->l_spec.as<uint4>(), e->l_states.as<uint4>(), so);
        LZB_DBG("k_long_replay");
        e->timer.mark(s);  // long_replay
        if (n_segl) {
            k_long_stitch_a<<<(n_rseg + 63) / 64, 64, 0, s>>>(src, src_off, src_len, bs, rseg, n_rseg, words, e->l_spec.as<uint4>(), e->l_states.as<uint4>(),
                                                               e->l_fix.as<uint4>(), so);
            LZB_DBG("k_long_stitch_a");
            k_long_stitch_b<<<(n_segl + 3) / 4, 128, 0, s>>>(src, src_off, src_len, st, bs, sl, n_segl, words, e->l_spec.as<uint4>(), e->l_states.as<uint4>(),
                                                              e->l_fix.as<uint4>(), so, e->l_tail.as<uint4>(), ctr + 13);
        }
        LZB_DBG("k_long_stitch_b");
        e->timer.mark(s);  // long_stitch
        if (n_segl) {
            k_long_seg_stats<<<(n_rseg + 3) / 4, 128, 0, s>>>(e->l_spec.as<uint4>(), e->l_fix.as<uint4>(), so, n_rseg, e->l_agg.as<SegAgg>(), e->l_csum.as<uint2>());
            LZB_DBG("k_long_seg_stats");
            static const int bw = getenv("LZB_BLOCKS_WARPS") ? atoi(getenv("LZB_BLOCKS_WARPS")) : 4;
            k_long_blocks<<<(n_segl + bw - 1) / bw, bw * 32, 0, s>>>(src_off, src_len, e->streams.as<EncStream>(), bs, sl, n_segl, e->l_spec.as<uint4>(), e->l_fix.as<uint4>(), so,
                                                            e->l_agg.as<SegAgg>(), e->l_csum.as<uint2>(), e->l_entry.as<SegEntry>(), e->l_tail.as<uint4>(), e->packs.as<uint2>(),
                                                            e->block_ids.as<uint32_t>(), e->blocks.as<EncBlock>(), ctr, ctr + 14);
            LZB_DBG("k_long_blocks");
            k_long_write_packs<<<(n_rseg + 3) / 4, 128, 0, s>>>(bs, rseg, n_rseg, e->l_spec.as<uint4>(), e->l_fix.as<uint4>(), so, e->l_agg.as<SegAgg>(),
                                                                 e->l_entry.as<SegEntry>(), e->packs.as<uint2>());
            e->launches += 6;
        }
        LZB_DBG("k_long_write_packs");
        e->timer.mark(s);  // long_packs
    }
    if (tot.n_literals) {
        unsigned g = (unsigned)((tot.n_literals + kFseEncWarps - 1) / kFseEncWarps);
        if (g > (unsigned)e->n_sms * 4) g = (unsigned)e->n_sms * 4;
        k_enc_fse_blocks<<<g, kFseEncWarps * 32, 0, s>>>(e->blocks.as<EncBlock>(), ctr, e->packs.as<uint2>(),
                                                                                        e->lits.as<uint8_t>(), e->out.as<uint8_t>(), ctr + 2, src);
        e->launches += 1;
    }
    e->timer.mark(s);  // fse_blocks
    k_enc_assemble<<<(unsigned)((n + kAsmWarps - 1) / kAsmWarps), kAsmWarps * 32, 0, s>>>(
        src, src_off, src_len, dst, dst_off, dst_cap, n, e->streams.as<EncStream>(), e->counts.as<StreamCounts>(), e->block_ids.as<uint32_t>(),
        e->blocks.as<EncBlock>(), e->out.as<uint8_t>(), out_len, status);
    e->launches += 1;
    if (n_long) {
        k_long_copy<<<(unsigned)e->n_sms * 8, 256, 0, s>>>(e->blocks.as<EncBlock>(), ctr, e->out.as<uint8_t>(), dst);
        e->launches += 1;
    }
    e->timer.mark(s);  // assemble
    LZB_CK(e, cudaGetLastError());
    if (!wait) {  // lzfse_b200_encoder_sync (or the caller's own wait on `s`) completes the call
        e->pending = true;
        e->pending_stream = s;
        return LZFSE_B200_OK;
    }
    LZB_CK(e, cudaStreamSynchronize(s));
    e->tables_dirty = false;
    e->timer.finish();
    return LZFSE_B200_OK;
}

}  // namespace

extern "C" {

int lzfse_b200_encoder_create(int device, lzfse_b200_encoder **out) {
    if (!out) return LZFSE_B200_INVALID_ARGUMENT;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count) { cudaGetLastError(); return LZFSE_B200_NO_DEVICE; }
    DeviceGuard g(device);
    if (!g.ok) return LZFSE_B200_NO_DEVICE;
    lzfse_b200_encoder *e = new (std::nothrow) lzfse_b200_encoder();
    if (!e) return LZFSE_B200_OUT_OF_MEMORY;
    e->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete e; return LZFSE_B200_NO_DEVICE; }
    e->n_sms = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&e->own_stream, cudaStreamNonBlocking) != cudaSuccess) { delete e; return LZFSE_B200_CUDA_ERROR; }
    if (const char *ef = getenv("LZB_ENC_FAST")) e->allow_fast = atoi(ef) != 0;
    if (const char *ef = getenv("LZB_ENC_LONG")) e->allow_long = atoi(ef) != 0;
    if (const char *ef = getenv("LZB_ENC_SEG")) e->allow_seg = atoi(ef) != 0;
    cudaFuncAttributes fa;
    if (cudaFuncGetAttributes(&fa, k_enc_fse_blocks) != cudaSuccess ||
        cudaFuncSetAttribute(k_enc_find, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFindSmemBytes) != cudaSuccess ||
        cudaFuncSetAttribute(k_long_chain, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kChainSmem) != cudaSuccess ||
        cudaFuncSetAttribute(k_long_heads, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kHeadsSmem) != cudaSuccess) {
        cudaGetLastError();  // no sm_100a image for this device: there is no fallback path
        cudaStreamDestroy(e->own_stream);
        delete e;
        return LZFSE_B200_NO_DEVICE;
    }
    *out = e;
    return LZFSE_B200_OK;
}

void lzfse_b200_encoder_destroy(lzfse_b200_encoder *e) {
    if (!e) return;
    DeviceGuard g(e->device);
    for (DevBuf *b : {&e->streams, &e->counts, &e->totals_dev, &e->tables, &e->packs, &e->lits, &e->block_ids, &e->blocks, &e->out, &e->counters, &e->words,
                      &e->long_list, &e->seg_list, &e->l_prev, &e->l_heads, &e->l_cseg, &e->l_rseg, &e->l_out, &e->l_spec, &e->l_fix, &e->l_states, &e->l_tail, &e->l_agg, &e->l_csum, &e->l_entry}) b->release();
    e->totals_host.release();
    e->stage.release();
    e->timer.release();
    if (e->own_stream) cudaStreamDestroy(e->own_stream);
    delete e;
}

const char *lzfse_b200_encoder_last_error(const lzfse_b200_encoder *e) { return e ? e->last_error.c_str() : ""; }
uint64_t lzfse_b200_encoder_last_launches(const lzfse_b200_encoder *e) { return e ? e->launches : 0; }
void lzfse_b200_encoder_set_timing(lzfse_b200_encoder *e, int enabled) { if (e) e->timer.enabled = enabled != 0; }
int lzfse_b200_encoder_last_stage_ms(const lzfse_b200_encoder *e, float *ms, int cap) {
    if (!e) return 0;
    for (int i = 0; i < e->timer.n_done && i < cap; i++) ms[i] = e->timer.ms[i];
    return e->timer.n_done;
}
// Test hook (not part of the public header): copies the per-position words of the last fast parse to the host.
size_t lzfse_b200_debug_encoder_words(lzfse_b200_encoder *e, uint32_t *host, size_t max_words) {
    if (!e || !e->words.p) return 0;
    DeviceGuard g(e->device);
    size_t n = e->words.cap / sizeof(uint32_t);
    if (n > max_words) n = max_words;
    if (cudaMemcpy(host, e->words.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost) != cudaSuccess) return 0;
    return n;
}
size_t lzfse_b200_debug_encoder_packs(lzfse_b200_encoder *e, uint64_t *host, size_t max_packs) {
    if (!e || !e->packs.p) return 0;
    DeviceGuard g(e->device);
    size_t n = e->packs.cap / sizeof(uint2);
    if (n > max_packs) n = max_packs;
    if (cudaMemcpy(host, e->packs.p, n * sizeof(uint2), cudaMemcpyDeviceToHost) != cudaSuccess) return 0;
    return n;
}
size_t lzfse_b200_encode_bound(size_t n) { return n + n / 4 + (n / 16384 + 2) * 768 + 64; }
size_t lzfse_b200_encode_bound_strict(size_t n) { return (size_t)out_cap(n) + 64; }

int lzfse_b200_encode_batch_device(lzfse_b200_encoder *e, const uint8_t *src, const uint64_t *src_off, const uint64_t *src_len, uint8_t *dst,
                                   const uint64_t *dst_off, const uint64_t *dst_cap, uint64_t *out_len, int32_t *status, size_t n, void *stream) {
    if (!e || (n && (!src_off || !src_len || !dst_off || !dst_cap || !out_len || !status))) return LZFSE_B200_INVALID_ARGUMENT;
    DeviceGuard g(e->device);
    if (!g.ok) return LZFSE_B200_CUDA_ERROR;
    return encode_batch_device_impl(e, src, src_off, src_len, dst, dst_off, dst_cap, out_len, status, n, (cudaStream_t)stream);
}

int lzfse_b200_encode_batch_device_async(lzfse_b200_encoder *e, const uint8_t *src, const uint64_t *src_off, const uint64_t *src_len, uint8_t *dst,
                                         const uint64_t *dst_off, const uint64_t *dst_cap, uint64_t *out_len, int32_t *status, size_t n, void *stream) {
    if (!e || (n && (!src_off || !src_len || !dst_off || !dst_cap || !out_len || !status))) return LZFSE_B200_INVALID_ARGUMENT;
    DeviceGuard g(e->device);
    if (!g.ok) return LZFSE_B200_CUDA_ERROR;
    return encode_batch_device_impl(e, src, src_off, src_len, dst, dst_off, dst_cap, out_len, status, n, (cudaStream_t)stream, false);
}

int lzfse_b200_encoder_sync(lzfse_b200_encoder *e) {
    if (!e) return LZFSE_B200_INVALID_ARGUMENT;
    if (!e->pending) return LZFSE_B200_OK;
    DeviceGuard g(e->device);
    if (!g.ok) return LZFSE_B200_CUDA_ERROR;
    e->pending = false;
    LZB_CK(e, cudaStreamSynchronize(e->pending_stream));
    e->tables_dirty = false;
    e->timer.finish();
    return LZFSE_B200_OK;
}

int lzfse_b200_encode_batch_host(lzfse_b200_encoder *e, const uint8_t *src, const uint64_t *src_off, const uint64_t *src_len, uint8_t *dst,
                                 const uint64_t *dst_off, const uint64_t *dst_cap, uint64_t *out_len, int32_t *status, size_t n) {
    if (!e || (n && (!src || !src_off || !src_len || !dst_off || !dst_cap || !out_len || !status))) return LZFSE_B200_INVALID_ARGUMENT;
    if (n == 0) return LZFSE_B200_OK;
    DeviceGuard g(e->device);
    if (!g.ok) return LZFSE_B200_CUDA_ERROR;
    cudaStream_t s = e->own_stream;
    HostStage &st = e->stage;
    int rc = stage_sources(e, st, src, src_off, src_len, n, 4 * n, s);
    if (rc) return rc;
    rc = stage_outputs(e, st, dst_off, dst_cap, n);
    if (rc) return rc;
    LZB_CK(e, st.desc.reserve(4 * n * sizeof(uint64_t)));
    LZB_CK(e, st.res.reserve(n * (sizeof(uint64_t) + sizeof(int32_t))));
    LZB_CK(e, cudaMemcpyAsync(st.desc.p, st.pin.p, 4 * n * sizeof(uint64_t), cudaMemcpyHostToDevice, s));
    uint64_t *dd = st.desc.as<uint64_t>();
    uint64_t *d_out_len = st.res.as<uint64_t>();
    int32_t *d_status = reinterpret_cast<int32_t *>(d_out_len + n);
    rc = encode_batch_device_impl(e, st.src.as<uint8_t>(), dd, dd + n, st.dst.as<uint8_t>(), dd + 2 * n, dd + 3 * n, d_out_len, d_status, n, s);
    if (rc) return rc;
    LZB_CK(e, cudaMemcpyAsync(out_len, d_out_len, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
    LZB_CK(e, cudaMemcpyAsync(status, d_status, n * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    LZB_CK(e, cudaStreamSynchronize(s));
    rc = fetch_outputs(e, st, dst, dst_off, out_len, status, n, s);
    if (rc) return rc;
    LZB_CK(e, cudaStreamSynchronize(s));
    return LZFSE_B200_OK;
}

int lzfse_b200_encode_bytes(lzfse_b200_encoder *e, const uint8_t *src, size_t src_len, uint8_t *dst, size_t dst_cap, size_t *dst_len) {
    if (!e || (!src && src_len) || (!dst && dst_cap)) return LZFSE_B200_INVALID_ARGUMENT;
    uint64_t so = 0, sl = src_len, doff = 0, dc = dst_cap, ol = 0;
    int32_t st = 0;
    static const uint8_t empty = 0;
    int rc = lzfse_b200_encode_batch_host(e, src ? src : &empty, &so, &sl, dst, &doff, &dc, &ol, &st, 1);
    if (dst_len) *dst_len = (size_t)ol;
    return rc ? rc : st;
}

}  // extern "C"
