// decode.cu -- batched LZFSE decode kernels for sm_100a.
//
// Pipeline (all launches on one CUDA stream, see api.cu):
//   k_scan<false>   thread / stream   walk block headers, count blocks / literals / LMDs, validate
//   k_exclusive_scan                   per-stream bases (k_scan_tile_sums / k_scan_tile_apply around it for > 8 Ki streams)
//   k_scan<true>    thread / stream   fill BlockDesc[] / FseDesc[]
//   k_fse_literals  lane / FSE block  weights -> U table in shared memory -> 4-state literal decode
//   k_fse_lmds      lane / FSE block  weights -> L/M/D table in shared memory -> LMD decode + validation
//   k_expand_vn     warp / stream     streams that are one small LZVN block: 32 payload bytes per step (only launched
//                                     when the batch has raw or LZVN blocks)
//   k_expand        warp / stream     literal placement + match copies, raw blocks, LZVN blocks inside longer frames
//                                     (or k_expand_cta, expand.cu, for batches of few large streams)
//   k_finish        thread / stream   error key -> status, out_len
//
// The entropy stages map one LANE to one block: an FSE stream is a serial chain (the bit position of
// symbol n+1 depends on symbol n), so the only parallelism is across blocks; a lane-per-block warp
// keeps all 32 lanes issuing, and the tables are laid out [state][lane] so a warp's 32 lookups hit
// 32 different banks.  What the reference does in Literals::load / FseCore::decode_internal
// (fse/literals.rs:49-91, fse/fse_core.rs:91-141) is restated below with every check it makes.
#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "lz_blocks.cuh"

namespace lzb {

// ------------------------------------------------------------------------------------------------
// Frame scan (decode/decoder.rs:72-99 dispatch loop, fse/block.rs:80-136,218-341 header parse + validate)
// ------------------------------------------------------------------------------------------------

struct FseHeader {
    uint32_t n_raw, n_literals, n_lit_payload, lit_bits, n_lmds, n_lmd_payload, lmd_bits, n_weight_bytes, header_size;
    uint16_t lit_state[4], lmd_state[3];
    bool v1;
};

__device__ int validate_fse_header(const FseHeader &h) {
    // LmdParam::validate (fse/block.rs:267-283)
    uint32_t lmd_limit = 1024 + 8 + (h.n_lmds * kMaxLBits + h.n_lmds * kMaxMBits + h.n_lmds * kMaxDBits + 7) / 8;
    if (h.n_lmds > kLmdsPerBlock || h.n_lmd_payload < 8 || h.n_lmd_payload > lmd_limit) return LZFSE_B200_FSE_BAD_LMD_COUNT;
    if (h.lmd_bits > 7) return LZFSE_B200_FSE_BAD_LMD_BITS;
    if (h.lmd_state[0] >= kLStates || h.lmd_state[1] >= kMStates || h.lmd_state[2] >= kDStates) return LZFSE_B200_FSE_BAD_LMD_STATE;
    // LiteralParam::validate (fse/block.rs:324-341)
    if (h.n_literals % 4 != 0 || h.n_literals > kLiteralsPerBlock) return LZFSE_B200_FSE_BAD_LITERAL_COUNT;
    if (h.n_lit_payload > 1024 + (h.n_literals * kMaxUBits + 7) / 8) return LZFSE_B200_FSE_BAD_LITERAL_COUNT;
    if (h.lit_bits > 7) return LZFSE_B200_FSE_BAD_LITERAL_BITS;
    if (h.lit_state[0] >= kUStates || h.lit_state[1] >= kUStates || h.lit_state[2] >= kUStates || h.lit_state[3] >= kUStates)
        return LZFSE_B200_FSE_BAD_LMD_PAYLOAD;  // sic: the reference reports BadLmdPayload here
    // FseBlock::validate (fse/block.rs:218-227)
    if (h.n_raw > h.n_literals + h.n_lmds * kMaxMValue) return LZFSE_B200_FSE_BAD_RAW_BYTE_COUNT;
    return LZFSE_B200_OK;
}

__device__ int parse_v2(const uint8_t *s, FseHeader &h) {
    h.v1 = false;
    h.n_raw = ld_u32(s + 4);
    uint64_t p = ld_u64(s + 8);
    h.n_literals = (uint32_t)(p & 0xFFFFF);
    h.n_lit_payload = (uint32_t)((p >> 20) & 0xFFFFF);
    h.n_lmds = (uint32_t)((p >> 40) & 0xFFFFF);
    h.lit_bits = 7 - (uint32_t)((p >> 60) & 7);
    p = ld_u64(s + 16);
    for (int i = 0; i < 4; i++) h.lit_state[i] = (uint16_t)((p >> (10 * i)) & 0x3FF);
    h.n_lmd_payload = (uint32_t)((p >> 40) & 0xFFFFF);
    h.lmd_bits = 7 - (uint32_t)((p >> 60) & 7);
    p = ld_u64(s + 24);
    uint32_t header_size = (uint32_t)p;
    h.lmd_state[0] = (uint16_t)((p >> 32) & 0x3FF);
    h.lmd_state[1] = (uint16_t)((p >> 42) & 0x3FF);
    h.lmd_state[2] = (uint16_t)((p >> 52) & 0x3FF);
    h.n_weight_bytes = header_size - kV2HeaderSize;
    h.header_size = header_size;
    if (h.n_weight_bytes > kV2WeightBytesMax) return LZFSE_B200_FSE_BAD_WEIGHT_PAYLOAD;
    return validate_fse_header(h);
}

__device__ int parse_v1(const uint8_t *s, FseHeader &h) {
    h.v1 = true;
    h.n_raw = ld_u32(s + 4);
    uint32_t n_payload = ld_u32(s + 8);
    h.n_literals = ld_u32(s + 12);
    h.n_lmds = ld_u32(s + 16);
    h.n_lit_payload = ld_u32(s + 20);
    h.n_lmd_payload = ld_u32(s + 24);
    h.lit_bits = 0u - ld_u32(s + 28);
    for (int i = 0; i < 4; i++) h.lit_state[i] = (uint16_t)ld_u16(s + 32 + 2 * i);
    h.lmd_bits = 0u - ld_u32(s + 40);
    for (int i = 0; i < 3; i++) h.lmd_state[i] = (uint16_t)ld_u16(s + 44 + 2 * i);
    h.n_weight_bytes = kV1WeightBytes;
    h.header_size = kV1HeaderSize + kV1WeightBytes;
    if (n_payload < h.n_lit_payload + h.n_lmd_payload) return LZFSE_B200_FSE_BAD_PAYLOAD_COUNT;
    return validate_fse_header(h);
}

template <bool FILL>
__global__ void k_scan(const uint8_t *__restrict__ src_base, const uint64_t *__restrict__ src_off,
                       const uint64_t *__restrict__ src_len, const uint64_t *__restrict__ dst_off,
                       const uint64_t *__restrict__ dst_cap, size_t n, StreamCounts *counts /* in FILL: scanned bases */,
                       BlockDesc *blocks, FseDesc *fse, uint32_t *err, uint64_t *raw_total, uint32_t *n_blocks_out,
                       uint32_t *work /* kWorkWords counters, or null (probe) */, uint32_t long_fse /* streams with at least this many bvx blocks are long */,
                       uint64_t *long_base, uint32_t *long_blocks, uint32_t *long_streams,
                       const uint64_t *__restrict__ limit /* bounded decode: stop once this many bytes are covered, or null */, uint8_t *more_out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bool more = false;
    const uint8_t *s = src_base + src_off[i];
    const uint64_t len = src_len[i];
    uint64_t pos = 0, raw = 0;
    uint32_t blk = 0, n_fse = 0;
    uint64_t n_lit = 0, n_lmd = 0;
    uint32_t key = kNoError;
    StreamCounts base;
    bool is_long = false;   // expanded by the two-pass kernels (expand_long.cu)
    uint32_t long_slot = 0;
    if (FILL) {
        base = counts[i];
        const StreamCounts next = counts[i + 1];
        is_long = next.n_fse - base.n_fse >= long_fse;
        if (is_long) {  // image range, slots in the block list, entry in the stream list
            long_base[i] = atomicAdd(reinterpret_cast<unsigned long long *>(work + 12), (unsigned long long)((raw_total[i] + 3) & ~3ull));
            long_slot = atomicAdd(work + 14, (uint32_t)(next.n_blocks - base.n_blocks));
            long_streams[atomicAdd(work + 15, 1u)] = (uint32_t)i;
        } else {
            long_base[i] = ~0ull;
        }
    }
    for (;;) {
        uint64_t rest = len - pos;
        uint32_t kb = blk < 0x1FFFFFu ? blk : 0x1FFFFFu;
        // Bounded decode (the reference's decode_n, fse_core.rs:143-198): whole blocks until the limit is covered.  What
        // follows is not looked at -- unless the frame ends exactly here, which is then checked like any other end.
        if (limit && raw >= limit[i]) {
            bool go_on = false;  // blocks that add nothing (and the end-of-stream marker) are still walked and checked
            if (raw == limit[i] && rest >= 4) {
                const uint32_t m = ld_u32(s + pos);
                go_on = m == kMagicEos || (rest >= 8 && (m == kMagicRaw || m == kMagicVxn || m == kMagicVx1 || m == kMagicVx2) && ld_u32(s + pos + 4) == 0);
            }
            if (!go_on) { more = true; break; }
        }
        if (rest < 4) { key = err_key(kb, PH_HEADER, LZFSE_B200_PAYLOAD_UNDERFLOW); break; }
        uint32_t magic = ld_u32(s + pos);
        if (magic == kMagicEos) {
            if (rest != 4) key = err_key(kb, PH_HEADER, LZFSE_B200_PAYLOAD_OVERFLOW);
            break;
        }
        BlockDesc bd;
        bd.src_off = src_off[i] + pos;
        bd.dst_off = (FILL ? dst_off[i] : 0) + raw;
        bd.stream = (uint32_t)i;
        bd.index = blk;
        bd.fse_idx = 0;
        bd.pad = 0;
        bool stop = false;
        uint64_t blen;
        if (magic == kMagicRaw) {
            if (rest < 8) { key = err_key(kb, PH_HEADER, LZFSE_B200_PAYLOAD_UNDERFLOW); break; }
            bd.n_raw = ld_u32(s + pos + 4);
            bd.type = BT_RAW;
            blen = 8ull + bd.n_raw;
            // raw/block.rs:71-93: copy what is there, then PayloadUnderflow if the input was short.  The
            // copy is where a too small dst shows up first (C-ABI BufferOverflow).
            const uint64_t avail = rest - 8 < bd.n_raw ? rest - 8 : bd.n_raw;
            if (dst_cap != nullptr && raw + avail > dst_cap[i]) { key = err_key(kb, PH_LMD, LZFSE_B200_BUFFER_OVERFLOW); break; }
            if (rest < blen) { key = err_key(kb, PH_LMD, LZFSE_B200_PAYLOAD_UNDERFLOW); break; }
        } else if (magic == kMagicVxn) {
            if (rest < kVnHeaderSize) { key = err_key(kb, PH_HEADER, LZFSE_B200_PAYLOAD_UNDERFLOW); break; }
            bd.n_raw = ld_u32(s + pos + 4);
            bd.type = BT_VXN;
            blen = (uint64_t)kVnHeaderSize + ld_u32(s + pos + 8);
            if (rest < blen) stop = true;  // the opcode interpreter decides which error this is
            // An LZVN header can announce any n_raw_bytes; the reference only compares it with what the opcodes
            // produced (vn_core.rs:96-111).  When the announcement would cross the 32-bit position limit below but the
            // destination is smaller than that anyway, the block cannot succeed: let the interpreter find the error
            // the reference finds (bad opcode, BufferOverflow at the write that does not fit, VnBadPayload at the
            // end) and keep the positions of the descriptors inside the destination.
            // Bounded decode sizes its internal buffer from what the blocks announce: an LZVN opcode of at most 3 bytes
            // produces at most 271, so a header that announces more than 136 bytes per payload byte cannot be honoured
            // (the interpreter will end with VnBadPayload); it only gets the room its payload could fill.
            if (limit && (uint64_t)bd.n_raw > 136ull * ld_u32(s + pos + 8) + 16) {
                stop = true;
                bd.n_raw = 136u * ld_u32(s + pos + 8) + 16;
                bd.pad = 1;
            }
            if (raw + bd.n_raw > kMaxStreamRaw && dst_cap != nullptr && dst_cap[i] < raw + bd.n_raw) {
                stop = true;
                bd.n_raw = (uint32_t)(dst_cap[i] > raw ? dst_cap[i] - raw : 0);
                bd.pad = 1;  // n_raw is not the header's: one-lane interpreter only (vn_fast_eligible)
            }
        } else if (magic == kMagicVx2 || magic == kMagicVx1) {
            const bool v1 = magic == kMagicVx1;
            const uint32_t hs = v1 ? kV1HeaderSize : kV2HeaderSize;
            if (rest < hs) { key = err_key(kb, PH_HEADER, LZFSE_B200_PAYLOAD_UNDERFLOW); break; }
            FseHeader h;
            int e = v1 ? parse_v1(s + pos, h) : parse_v2(s + pos, h);
            if (e) { key = err_key(kb, PH_HEADER, e); break; }
            if (rest - hs < h.n_weight_bytes) { key = err_key(kb, PH_HEADER, LZFSE_B200_PAYLOAD_UNDERFLOW); break; }
            bd.n_raw = h.n_raw;
            bd.type = v1 ? BT_VX1 : BT_VX2;
            bd.fse_idx = (uint32_t)(FILL ? base.n_fse : 0) + n_fse;
            uint32_t flags = v1 ? FSE_V1 : 0;
            // fse_core.rs:62-88: each payload is `take`n after the previous stage succeeded.
            if (rest < (uint64_t)h.header_size + h.n_lit_payload) {
                flags |= FSE_TRUNC_LIT; stop = true;
                key = err_key(kb, PH_LIT_TAKE, LZFSE_B200_PAYLOAD_UNDERFLOW);
            } else if (rest < (uint64_t)h.header_size + h.n_lit_payload + h.n_lmd_payload) {
                flags |= FSE_TRUNC_LMD; stop = true;
                key = err_key(kb, PH_LMD_TAKE, LZFSE_B200_PAYLOAD_UNDERFLOW);
            }
            blen = (uint64_t)h.header_size + h.n_lit_payload + h.n_lmd_payload;
            if (FILL) {
                FseDesc fd;
                fd.lit_off = base.n_literals + n_lit;
                fd.lmd_off = base.n_lmds + n_lmd;
                fd.block = (uint32_t)base.n_blocks + blk;
                fd.flags = flags;
                fd.header_size = h.header_size;
                fd.n_weight_bytes = h.n_weight_bytes;
                fd.n_literals = h.n_literals; fd.n_lit_payload = h.n_lit_payload; fd.lit_bits = h.lit_bits;
                fd.n_lmds = h.n_lmds; fd.n_lmd_payload = h.n_lmd_payload; fd.lmd_bits = h.lmd_bits;
                for (int k = 0; k < 4; k++) fd.lit_state[k] = h.lit_state[k];
                for (int k = 0; k < 3; k++) fd.lmd_state[k] = h.lmd_state[k];
                fd.pad = 0; fd.pad2 = 0;
                fd.n_raw = h.n_raw;
                fd.ok_lit = 0; fd.ok_lmd = 0;
                fse[bd.fse_idx] = fd;
            }
            n_fse++;
            n_lit += (h.n_literals + 15) & ~15u;  // 16-byte aligned literal runs
            n_lmd += (h.n_lmds + 1) & ~1u;        // 16-byte aligned LMD runs
        } else {
            key = err_key(kb, PH_HEADER, LZFSE_B200_BAD_BLOCK);
            break;
        }
        // C-ABI limit: the expansion stage keeps stream positions in 32 bits
        if (raw + bd.n_raw > kMaxStreamRaw) { key = err_key(kb, PH_HEADER, LZFSE_B200_BUFFER_OVERFLOW); break; }
        if (FILL) {
            blocks[base.n_blocks + blk] = bd;
            if (is_long) long_blocks[long_slot + blk] = (uint32_t)(base.n_blocks + blk);
        }
        raw += bd.n_raw;
        blk++;
        pos += blen;
        if (stop) break;
    }
    if (!FILL) {
        StreamCounts c;
        c.n_blocks = blk; c.n_fse = n_fse; c.n_literals = n_lit; c.n_lmds = n_lmd;
        counts[i] = c;
        err[i] = key;
        raw_total[i] = raw;
        if (n_blocks_out) n_blocks_out[i] = blk;
        if (more_out) more_out[i] = more ? 1 : 0;
        if (work && n_fse >= long_fse) {  // totals of the long streams: the host sizes their image from these
            atomicAdd(reinterpret_cast<unsigned long long *>(work + 8), (unsigned long long)((raw + 3) & ~3ull));
            atomicAdd(work + 10, blk);
            atomicAdd(work + 11, 1u);
        }
    }
}

// Exclusive scan of StreamCounts in place with one CTA; totals[0] receives the grand totals and
// counts[n] too (so stream i's range is [counts[i], counts[i+1])).
__device__ __forceinline__ StreamCounts sc_add(StreamCounts a, const StreamCounts &b) {
    a.n_blocks += b.n_blocks; a.n_fse += b.n_fse; a.n_literals += b.n_literals; a.n_lmds += b.n_lmds;
    return a;
}
__device__ __forceinline__ StreamCounts sc_shfl_up(const StreamCounts &v, int o) {
    StreamCounts r;
    r.n_blocks = __shfl_up_sync(0xFFFFFFFFu, v.n_blocks, o); r.n_fse = __shfl_up_sync(0xFFFFFFFFu, v.n_fse, o);
    r.n_literals = __shfl_up_sync(0xFFFFFFFFu, v.n_literals, o); r.n_lmds = __shfl_up_sync(0xFFFFFFFFu, v.n_lmds, o);
    return r;
}
__device__ __forceinline__ StreamCounts sc_warp_inclusive(StreamCounts v, uint32_t lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const StreamCounts t = sc_shfl_up(v, o);
        if (lane >= (uint32_t)o) v = sc_add(v, t);
    }
    return v;
}
__global__ void __launch_bounds__(1024) k_exclusive_scan(StreamCounts *counts, size_t n, StreamCounts *totals, const uint32_t *work) {
    __shared__ StreamCounts warp_tot[32];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t per = (n + blockDim.x - 1) / blockDim.x;
    const size_t lo = (size_t)threadIdx.x * per, hi = lo + per < n ? lo + per : n;
    const uint4 *c4 = reinterpret_cast<const uint4 *>(counts);
    StreamCounts acc = {0, 0, 0, 0};
#pragma unroll 4
    for (size_t i = lo; i < hi; i++) {  // independent 16-byte loads: the unrolled body keeps eight of them in flight
        const uint4 a = c4[2 * i], b = c4[2 * i + 1];
        acc.n_blocks += a.x | ((uint64_t)a.y << 32); acc.n_fse += a.z | ((uint64_t)a.w << 32);
        acc.n_literals += b.x | ((uint64_t)b.y << 32); acc.n_lmds += b.z | ((uint64_t)b.w << 32);
    }
    // block-wide exclusive scan of the 1024 partial sums: warp scans through shuffles, then the 32 warp totals
    const StreamCounts inc = sc_warp_inclusive(acc, lane);
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const StreamCounts w = sc_warp_inclusive(warp_tot[lane], lane);
        if (lane == 31) { *totals = w; counts[n] = w; }
        if (work && lane < 4) reinterpret_cast<uint32_t *>(totals + 1)[lane] = work[8 + lane];  // LongTotals (see api.cu)
        StreamCounts ex = sc_shfl_up(w, 1);
        if (lane == 0) ex = StreamCounts{0, 0, 0, 0};
        warp_tot[lane] = ex;
    }
    __syncthreads();
    StreamCounts run = sc_add(warp_tot[warp], inc);
    run.n_blocks -= acc.n_blocks; run.n_fse -= acc.n_fse; run.n_literals -= acc.n_literals; run.n_lmds -= acc.n_lmds;  // exclusive
#pragma unroll 4
    for (size_t i = lo; i < hi; i++) {
        const StreamCounts v = counts[i];
        counts[i] = run;
        run = sc_add(run, v);
    }
}

// Batches of many streams: the same scan in three short launches over 1024-element tiles (tile sums, scan of the tile
// sums with the kernel above, tile-local scans + offsets).  One CTA walking 512 Ki stream records on its own took
// longer than the header scan it serves.
__global__ void __launch_bounds__(1024) k_scan_tile_sums(const StreamCounts *__restrict__ counts, size_t n, StreamCounts *tile_sums) {
    __shared__ StreamCounts warp_tot[32];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t i = (size_t)blockIdx.x * 1024 + threadIdx.x;
    StreamCounts v = {0, 0, 0, 0};
    if (i < n) v = counts[i];
    const StreamCounts inc = sc_warp_inclusive(v, lane);
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const StreamCounts w = sc_warp_inclusive(warp_tot[lane], lane);
        if (lane == 31) tile_sums[blockIdx.x] = w;
    }
}
__global__ void __launch_bounds__(1024) k_scan_tile_apply(StreamCounts *counts, size_t n, const StreamCounts *__restrict__ tile_offsets, size_t n_tiles) {
    __shared__ StreamCounts warp_tot[32];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t i = (size_t)blockIdx.x * 1024 + threadIdx.x;
    StreamCounts v = {0, 0, 0, 0};
    if (i < n) v = counts[i];
    const StreamCounts inc = sc_warp_inclusive(v, lane);
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const StreamCounts w = sc_warp_inclusive(warp_tot[lane], lane);
        StreamCounts ex = sc_shfl_up(w, 1);
        if (lane == 0) ex = StreamCounts{0, 0, 0, 0};
        warp_tot[lane] = ex;
    }
    __syncthreads();
    StreamCounts run = sc_add(sc_add(tile_offsets[blockIdx.x], warp_tot[warp]), inc);
    run.n_blocks -= v.n_blocks; run.n_fse -= v.n_fse; run.n_literals -= v.n_literals; run.n_lmds -= v.n_lmds;  // exclusive
    if (i < n) counts[i] = run;
    if (i == n) counts[n] = tile_offsets[n_tiles];  // grand totals (the tile holding index n exists: see the launcher)
}
// counts must have room for n + 1 + (n / 1024 + 2) elements.
void launch_exclusive_scan(StreamCounts *counts, size_t n, StreamCounts *totals, const uint32_t *work, cudaStream_t s) {
    if (n <= 8192) {
        k_exclusive_scan<<<1, 1024, 0, s>>>(counts, n, totals, work);
        return;
    }
    const size_t n_tiles = (n + 1 + 1023) / 1024;  // covers index n as well
    StreamCounts *tiles = counts + n + 1;
    k_scan_tile_sums<<<(unsigned)n_tiles, 1024, 0, s>>>(counts, n, tiles);
    k_exclusive_scan<<<1, 1024, 0, s>>>(tiles, n_tiles, totals, work);
    k_scan_tile_apply<<<(unsigned)n_tiles, 1024, 0, s>>>(counts, n, tiles, n_tiles);
}

// ------------------------------------------------------------------------------------------------
// Weight payload reader (fse/weights.rs:66-105, fse/weight_encoder.rs:10-20)
// ------------------------------------------------------------------------------------------------
struct WeightReader {
    const uint8_t *p;
    uint32_t len, i;
    uint64_t accum;
    int accum_bits;
    uint32_t T, T_last;  // payload bits consumed so far / before the most recent weight
    bool v1;
    __device__ void init(const uint8_t *ptr, uint32_t n, bool is_v1) { p = ptr; len = n; i = 0; accum = 0; accum_bits = 0; T = 0; T_last = 0; v1 = is_v1; }
    __device__ uint32_t next() {
        if (v1) { uint32_t w = ld_u16(p + 2 * i); i++; return w; }
        // The reference tops its accumulator up one byte at a time while it holds <= 24 bits; a byte load per weight byte,
        // each lane in its own cache line, was a sixth of the literal kernel's stall samples.  Four bytes per refill (one
        // aligned word pair) decode the same weights: past the payload both read zeros.  What differs is how many bytes
        // have been pulled in at the end, and finish() reconstructs the reference's count from the bits consumed.
        if (accum_bits <= 32 && i < len) {
            const uint32_t take = len - i < 4u ? len - i : 4u;
            const uintptr_t a = reinterpret_cast<uintptr_t>(p + i);
            const uint32_t r = (uint32_t)a & 3u;
            const uint32_t *q = reinterpret_cast<const uint32_t *>(a - r);
            const uint32_t lo = __ldg(q), hi = (r && r + take > 4u) ? __ldg(q + 1) : 0u;  // only words that hold a payload byte
            uint32_t w = __funnelshift_r(lo, hi, r * 8);
            if (take < 4u) w &= (1u << (8 * take)) - 1u;
            accum |= (uint64_t)w << accum_bits;
            accum_bits += 8 * (int)take;
            i += take;
        }
        uint32_t u = (uint32_t)accum;
        uint32_t lo = u & 0x1F, bits, w;
        // WEIGHTS_BITS_TABLE / WEIGHTS_VALUE_TABLE (fse/constants.rs:115-124) in closed form
        if ((lo & 1) == 0) { bits = 2; w = (lo >> 1) & 1; }                 // x0: 00 -> 0, 10 -> 1
        else if ((lo & 3) == 1) { bits = 3; w = 2 + ((lo >> 2) & 1); }      // 001 -> 2, 101 -> 3
        else if ((lo & 7) == 3) { bits = 5; w = 4 + ((lo >> 3) & 3); }      // xx011 -> 4..7
        else if ((lo & 15) == 7) { bits = 8; w = 8 + ((u >> 4) & 0xF); }    // xxxx0111
        else { bits = 14; w = 24 + ((u >> 4) & 0x3FF); }                    // xxxxxxxxxx1111
        accum >>= bits;
        accum_bits -= (int)bits;
        T_last = T;
        T += bits;
        return w;
    }
    // Weights::load_v2 tail checks (fse/weights.rs:98-103) on the state the reference's byte-wise reader would be in: before
    // a weight it has pulled in bytes until it holds more than 24 bits (or the payload ends), so after the last weight it
    // has read min(len, (T_last + 24) / 8 + 1) bytes and holds that many bits minus what all weights consumed.
    __device__ int finish() const {
        if (v1) return LZFSE_B200_OK;
        uint32_t i_ref = (T_last + 24) / 8 + 1;
        if (i_ref > len) i_ref = len;
        const int bits_left = 8 * (int)i_ref - (int)T;
        if (bits_left < 0) return LZFSE_B200_FSE_WEIGHT_PAYLOAD_UNDERFLOW;
        if (bits_left >= 8 || i_ref != len) return LZFSE_B200_FSE_WEIGHT_PAYLOAD_OVERFLOW;
        return LZFSE_B200_OK;
    }
};

// check_totals (weights.rs:189-201)
__device__ __forceinline__ int check_weight_totals(uint32_t tl, uint32_t tm, uint32_t td, uint32_t tu) {
    return (tl <= kLStates && tm <= kMStates && td <= kDStates && tu <= kUStates) ? LZFSE_B200_OK : LZFSE_B200_FSE_BAD_WEIGHT_PAYLOAD;
}

// ------------------------------------------------------------------------------------------------
// Backward bit reader over global memory (bits/bit_reader.rs:20-71, bits/bit_src.rs:35-46).
//
// Like the reference, the reader keeps a bit cursor P (= 8*idx + accum_bits there) and re-reads a 64-bit
// window ending at P at every flush point: window bits [8*floor((P-57)/8), +64) always hold the next
// 57 bits.  The read is three aligned 32-bit loads plus two funnel shifts, branch-free, so the lanes of
// a warp (each at its own bit position) never diverge, and the loads are issued together with the
// table lookups of the same iteration.  `dead` reproduces "reads below index 0 yield 0": at a flush
// point the reference's idx is negative exactly when P < 57.
// ------------------------------------------------------------------------------------------------
struct BitWindow {
    const uint8_t *base;  // slice start
    uintptr_t pf_lo;      // lowest address worth prefetching (inside the block)
    int P;                // bit cursor, relative to the slice start
    bool dead;

    // slice = [start, start+len), len >= 8; `off` = unused high bits of the last byte.
    __device__ __forceinline__ int init(const uint8_t *start, uint32_t len, uint32_t off) {
        base = start; dead = false;
        pf_lo = reinterpret_cast<uintptr_t>(start);
        P = (int)len * 8 - (int)off;
        uint32_t last = start[len - 1];  // BitReader::new: the `off` bits above the cursor must be zero
        if (off != 0 && (last >> (8 - off)) != 0) return LZFSE_B200_BAD_BITSTREAM;
        return LZFSE_B200_OK;
    }
    // Flush point: returns the window; `cur` = position of the cursor inside it (57..64).
    __device__ __forceinline__ uint64_t window(int &cur) {
        if (P < 57) dead = true;
        const int byte = (P - 57) >> 3;  // the reference's idx: the window's top byte holds bit P-1
        cur = P - byte * 8;              // 57..64
        if (dead) return 0;
        const uintptr_t a = reinterpret_cast<uintptr_t>(base) + (intptr_t)byte;
        const uint32_t r = (uint32_t)a & 3u;
        const uint32_t *a4 = reinterpret_cast<const uint32_t *>(a - r);
        // Each lane walks its own stream, so one lane's L1 miss stalls the whole warp: keep every lane's
        // next lines resident by prefetching 2 lines below the window (no-op when already cached).
        const uintptr_t pf = a - r - 256;
        asm volatile("prefetch.global.L1 [%0];" ::"l"(pf > pf_lo ? pf : pf_lo));
        const uint32_t w0 = __ldg(a4), w1 = __ldg(a4 + 1);
        const uint32_t w2 = r ? __ldg(a4 + 2) : 0u;  // never touches a word that holds no slice byte
        const uint32_t lo = __funnelshift_r(w0, w1, r * 8), hi = __funnelshift_r(w1, w2, r * 8);
        return ((uint64_t)hi << 32) | lo;
    }
    // Same window without the dead-reader handling and the prefetch (callers guarantee P >= 57).
    __device__ __forceinline__ uint64_t window_fast(int &cur) const {
        const int byte = (P - 57) >> 3;
        cur = P - byte * 8;
        const uintptr_t a = reinterpret_cast<uintptr_t>(base) + (intptr_t)byte;
        const uint32_t r = (uint32_t)a & 3u;
        const uint32_t *a4 = reinterpret_cast<const uint32_t *>(a - r);
        const uint32_t w0 = __ldg(a4), w1 = __ldg(a4 + 1);
        const uint32_t w2 = r ? __ldg(a4 + 2) : 0u;
        const uint32_t lo = __funnelshift_r(w0, w1, r * 8), hi = __funnelshift_r(w1, w2, r * 8);
        return ((uint64_t)hi << 32) | lo;
    }
    __device__ __forceinline__ void prefetch() const {
        const uintptr_t pf = reinterpret_cast<uintptr_t>(base) + (intptr_t)((P >> 3) - 256);
        asm volatile("prefetch.global.L1 [%0];" ::"l"(pf > pf_lo ? pf : pf_lo));
    }
    __device__ __forceinline__ bool underflow() const { return P < 64; }  // BitReader::finalize
};
__device__ __forceinline__ uint32_t bfe(uint32_t v, uint32_t pos, uint32_t len) {
    uint32_t r;
    asm("bfe.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(v), "r"(pos), "r"(len));
    return r;
}
__device__ __forceinline__ uint32_t bits_at(uint64_t win, int pos, uint32_t n) {  // n < 32
    // one 64-bit funnel shift + one SGXT.W (szext: keep the low n bits, 0 for n == 0); bfe.u32 with a register length
    // costs five instructions on sm_100, shl + and-not two
    uint32_t r;
    asm("szext.wrap.u32 %0, %1, %2;" : "=r"(r) : "r"((uint32_t)(win >> pos)), "r"(n));
    return r;
}

// ------------------------------------------------------------------------------------------------
// The same reader, fed from shared memory.
//
// A divergent global load (32 lanes, 32 different lines) costs one L1 wavefront per lane, and three of
// them per symbol were more than half of the entropy kernels' time.  Each lane therefore owns a
// 128-byte ring in shared memory that mirrors the 128 stream bytes around its cursor (ring offset =
// global address mod 128); `cp.async` refills one 16-byte chunk at a time about 100 bytes ahead of the
// cursor, so the window is three shared-memory loads and the global traffic is asynchronous.
// ------------------------------------------------------------------------------------------------
constexpr uint32_t kRing = 128;                 // bytes mirrored per lane
constexpr uint32_t kRingStride = kRing + 16;    // per-lane stride (keeps 16-byte alignment, skews banks)
constexpr uint32_t kRingBytesPerWarp = 32 * kRingStride;

struct RingWindow {
    // All positions are 32-bit byte offsets from g0, a 16-byte aligned address 256 bytes below the stream's
    // first chunk (never dereferenced below lo16), so the upkeep is 32-bit arithmetic.
    uintptr_t g0;
    uint32_t a_base;         // offset of the slice start
    uint32_t lo16, hi;       // chunks below lo16 are never needed (zero-filled); bytes at or above hi are not read
    uint32_t next_chunk;     // offset of the next (lower) 16-byte chunk to fetch
    uint32_t ring;           // shared-memory address of this lane's ring
    int P;                   // bit cursor, relative to the slice start
    bool dead;

    __device__ __forceinline__ void fetch_chunk(uint32_t c) const {
        uint32_t n = 16, from = c;
        if (c - lo16 > hi - 16 - lo16) {  // not entirely inside [lo16, hi): partial or nothing (unsigned range check)
            n = c < lo16 ? 0u : (c < hi ? hi - c : 0u);
            if (n == 0) from = lo16;      // any valid address when nothing is read
        }
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(ring + (c & (kRing - 1))), "l"(g0 + from), "r"(n) : "memory");
    }
    // slice = [start, start+len), len >= 8; `off` = unused high bits of the last byte; [rlo, rhi) = the stream's bytes.
    __device__ __forceinline__ int init(const uint8_t *start, uint32_t len, uint32_t off, const uint8_t *rlo, const uint8_t *rhi,
                                        uint32_t ring_addr) {
        dead = false; ring = ring_addr;
        g0 = (reinterpret_cast<uintptr_t>(rlo) & ~(uintptr_t)15) - 256;
        lo16 = (uint32_t)(((reinterpret_cast<uintptr_t>(rlo) + 15) & ~(uintptr_t)15) - g0);
        hi = (uint32_t)(reinterpret_cast<uintptr_t>(rhi) - g0);
        a_base = (uint32_t)(reinterpret_cast<uintptr_t>(start) - g0);
        P = (int)len * 8 - (int)off;
        const uint32_t last = start[len - 1];  // BitReader::new: the `off` bits above the cursor must be zero
        if (off != 0 && (last >> (8 - off)) != 0) return LZFSE_B200_BAD_BITSTREAM;
        // first window: word-aligned offset a4 .. a4 + 12; fill the 8 chunks ending with the one that holds a4 + 11
        const uint32_t a = a_base + (uint32_t)((P - 57) >> 3);
        const uint32_t top = ((a & ~3u) + 11) & ~15u;
        for (uint32_t k = 0; k < kRing / 16; k++) fetch_chunk(top - 16 * k);
        next_chunk = top - kRing;
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        return LZFSE_B200_OK;
    }
    // Keeps the ring topped up: at most one chunk per call, so call it at least once per 16 bytes consumed.
    // A chunk is requested at least 100 bytes before the window reaches it, and the window cannot move 16 bytes
    // without another request being issued: by the time a chunk is read, six younger requests exist.  Waiting
    // until at most five requests are pending is therefore sufficient -- and it matters: waiting for all but the
    // newest one (wait_group 1) was a third of the entropy kernels' time in the profile (profiles/, r1b).
    __device__ __forceinline__ void refill() {
        const uint32_t a4 = (a_base + (uint32_t)((P - 57) >> 3)) & ~3u;
        if (a4 + 12 <= next_chunk + kRing) {  // the chunk slot above the window is free
            fetch_chunk(next_chunk);
            next_chunk -= 16;
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 5;" ::: "memory");
        }
    }
    // Window for the current cursor (callers guarantee P >= 57 and call refill() often enough).
    __device__ __forceinline__ uint64_t window_fast(int &cur) const {
        const int byte = (P - 57) >> 3;
        cur = P - byte * 8;
        const uint32_t a = a_base + (uint32_t)byte;
        const uint32_t r = a & 3u, o = (a - r) & (kRing - 1);
        uint32_t w0, w1, w2;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w0) : "r"(ring + o));
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w1) : "r"(ring + ((o + 4) & (kRing - 1))));
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w2) : "r"(ring + ((o + 8) & (kRing - 1))));
        const uint32_t lo = __funnelshift_r(w0, w1, r * 8), hi32 = __funnelshift_r(w1, w2, r * 8);
        return ((uint64_t)hi32 << 32) | lo;
    }
    // Flush point with the reference's dead-reader semantics (reads below index 0 yield 0).
    __device__ __forceinline__ uint64_t window(int &cur) {
        if (P < 57) dead = true;
        if (dead) { cur = P - ((P - 57) >> 3) * 8; return 0; }
        refill();
        return window_fast(cur);
    }
    __device__ __forceinline__ bool underflow() const { return P < 64; }  // BitReader::finalize
};

// ------------------------------------------------------------------------------------------------
// Literal stage: lane per block.  U table [1024][32] in shared memory, split into a 16-bit
// (k << 12 | delta) plane and an 8-bit symbol plane: 3 KiB per lane, 96 KiB per warp, 2 warps per SM.
// (fse/decoder.rs:299-335 build_u_table, fse/literals.rs:49-91 Literals::load)
// ------------------------------------------------------------------------------------------------
constexpr int kLitWarps = 2;
constexpr size_t kLitSmemPerWarp = 1024 * 32 * 3 + kRingBytesPerWarp;

__global__ void __launch_bounds__(kLitWarps * 32, 1)
k_fse_literals(const uint8_t *__restrict__ src_base, const uint64_t *__restrict__ src_off, const uint64_t *__restrict__ src_len,
               const BlockDesc *__restrict__ blocks, FseDesc *__restrict__ fse, uint32_t n_fse, uint8_t *__restrict__ lit_scratch,
               uint32_t *err, uint32_t *work_counter) {
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
    uint16_t *kd = reinterpret_cast<uint16_t *>(smem + warp * kLitSmemPerWarp);
    uint8_t *sy = smem + warp * kLitSmemPerWarp + 1024 * 32 * 2;
    const uint32_t ring_addr = (uint32_t)__cvta_generic_to_shared(smem + warp * kLitSmemPerWarp + 1024 * 32 * 3) + lane * kRingStride;
    for (;;) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(work_counter, 32u);
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (base >= n_fse) break;
        const uint32_t f = base + lane;
        if (f < n_fse) {
            FseDesc fd = fse[f];
            const BlockDesc bd = blocks[fd.block];
            const uint8_t *blk = src_base + bd.src_off;
            const bool v1 = fd.flags & FSE_V1;
            const uint8_t *wp = blk + (v1 ? kV1HeaderSize : kV2HeaderSize);
            const uint32_t kb = bd.index < 0x1FFFFFu ? bd.index : 0x1FFFFFu;
            // One pass over the weight payload: L/M/D weights are only summed here, the literal weights build the U table as
            // they are read (build_u_table, one lane per block, states in order); the payload's bit accounting and the totals
            // are checked afterwards, as the reference does before it builds anything -- a payload that fails has at worst
            // filled this lane's own table with nonsense (writes are clipped to it).
            WeightReader r;
            r.init(wp, fd.n_weight_bytes, v1);
            uint32_t tl = 0, tm = 0, td = 0, total = 0;
            for (int k = 0; k < 20; k++) tl += r.next();
            for (int k = 0; k < 20; k++) tm += r.next();
            for (int k = 0; k < 64; k++) td += r.next();
            for (uint32_t sym = 0; sym < 256; sym++) {
                uint32_t w = r.next();
                if (w == 0) continue;
                uint32_t k = __clz(w) - 21;  // clz(w) - clz(1024)
                uint32_t x = (2048u >> k) - w;
                const uint32_t room = total < 1024u ? 1024u - total : 0u;
                const uint32_t wn = w < room ? w : room;
                for (uint32_t j = 0; j < wn; j++) {
                    uint32_t kk, delta;
                    if (j < x) { kk = k; delta = ((w + j) << k) - 1024u; }
                    else { kk = k - 1; delta = (j - x) << (k - 1); }
                    kd[(total + j) * 32 + lane] = (uint16_t)(delta | (kk << 12));
                    sy[(total + j) * 32 + lane] = (uint8_t)sym;
                }
                total += w;
            }
            int e = r.finish();
            if (!e) e = check_weight_totals(tl, tm, td, total);
            if (e) {
                atomicMin(&err[bd.stream], err_key(kb, PH_WEIGHTS, e));
            } else if (!(fd.flags & FSE_TRUNC_LIT)) {
                for (uint32_t t = total; t < 1024; t++) { kd[t * 32 + lane] = (uint16_t)t; sy[t * 32 + lane] = 0; }

                // Literals::load.  The slice borrows the 8 bytes before the payload as the BitSrc pad
                // (fse_core.rs:30-33): it is never consumed by a well-formed stream.
                const uint8_t *s_lo = src_base + src_off[bd.stream], *s_hi = s_lo + src_len[bd.stream];
                RingWindow br;
                int st = br.init(blk + fd.header_size - 8, fd.n_lit_payload + 8, fd.lit_bits, s_lo, s_hi, ring_addr);
                if (st) {
                    atomicMin(&err[bd.stream], err_key(kb, PH_LIT, st));
                } else {
                    uint32_t s0 = fd.lit_state[0], s1 = fd.lit_state[1], s2 = fd.lit_state[2], s3 = fd.lit_state[3];
                    uint32_t *out = reinterpret_cast<uint32_t *>(lit_scratch + fd.lit_off);  // 16-byte aligned
                    const uint32_t n_it = fd.n_literals >> 2;
                    // One step = the reference's loop body: 4 literals, states 0..3 in that order, one flush.
                    // FAST = no dead-reader handling (the caller keeps P >= 57) and no per-step prefetch.
                    const uint16_t *kdl = kd + lane;  // this lane's column of the two table planes: one multiply-add per lookup
                    const uint8_t *syl = sy + lane;
                    auto step = [&](auto fast_tag) -> uint32_t {
                        constexpr bool FAST = decltype(fast_tag)::value;
                        int cur;
                        const uint64_t win = FAST ? br.window_fast(cur) : br.window(cur);
                        const uint32_t e0 = kdl[s0 * 32], e1 = kdl[s1 * 32], e2 = kdl[s2 * 32], e3 = kdl[s3 * 32];
                        const uint32_t y0 = syl[s0 * 32], y1 = syl[s1 * 32], y2 = syl[s2 * 32], y3 = syl[s3 * 32];
                        const uint32_t k0 = e0 >> 12, k1 = e1 >> 12, k2 = e2 >> 12, k3 = e3 >> 12;
                        const int p0 = cur - (int)k0, p1 = p0 - (int)k1, p2 = p1 - (int)k2, p3 = p2 - (int)k3;
                        s0 = bits_at(win, p0, k0) + (e0 & 0xFFF);
                        s1 = bits_at(win, p1, k1) + (e1 & 0xFFF);
                        s2 = bits_at(win, p2, k2) + (e2 & 0xFFF);
                        s3 = bits_at(win, p3, k3) + (e3 & 0xFFF);
                        br.P -= cur - p3;
                        return y0 | (y1 << 8) | (y2 << 16) | (y3 << 24);
                    };
                    uint32_t it = 0;
                    for (; it + 4 <= n_it && br.P >= 57 + 3 * 40; it += 4) {  // 16 literals, one 16-byte store
                        uint4 v;
                        br.refill();  // two steps consume at most 10 bytes
                        v.x = step(std::true_type{}); v.y = step(std::true_type{});
                        br.refill();
                        v.z = step(std::true_type{}); v.w = step(std::true_type{});
                        __stcg(reinterpret_cast<uint4 *>(out + it), v);
                    }
                    for (; it < n_it && br.P >= 57; it++) { br.refill(); out[it] = step(std::true_type{}); }
                    if (it != n_it) {  // the reader came within 57 bits of the pad: redo with the reference's exact flush semantics
                        br.init(blk + fd.header_size - 8, fd.n_lit_payload + 8, fd.lit_bits, s_lo, s_hi, ring_addr);
                        s0 = fd.lit_state[0]; s1 = fd.lit_state[1]; s2 = fd.lit_state[2]; s3 = fd.lit_state[3];
                        for (it = 0; it < n_it; it++) out[it] = step(std::false_type{});
                    }
                    { int cur; br.window(cur); }  // the final flush
                    if (br.underflow()) atomicMin(&err[bd.stream], err_key(kb, PH_LIT, LZFSE_B200_PAYLOAD_UNDERFLOW));
                    else if (s0 | s1 | s2 | s3) atomicMin(&err[bd.stream], err_key(kb, PH_LIT, LZFSE_B200_FSE_BAD_LMD_PAYLOAD));
                    else fse[f].ok_lit = 1;
                }
            }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// LMD stage: lane per block.  L/M/D table [384][32] x u32 in shared memory (48 KiB per warp).
// Entry: delta[0:8] (relative to the symbol kind's first state) | k[8:12] | v_bits[12:16] | [16:32] = v_base for
// L and M (<= 312), and 4 + (s & 3) for D, whose v_base = ((4 + (s & 3)) << (s >> 2)) - 4 = (field << v_bits) - 4.
// (fse/decoder.rs:244-292 build_v_table_block, fse/fse_core.rs:91-141 decode_internal)
// ------------------------------------------------------------------------------------------------
constexpr int kLmdWarps = 4;
constexpr size_t kLmdSmemPerWarp = 384 * 32 * 4 + kRingBytesPerWarp;

__device__ __forceinline__ uint32_t l_extra(uint32_t s) { return s < 16 ? 0u : (s == 16 ? 2u : (s == 17 ? 3u : (s == 18 ? 5u : 8u))); }
__device__ __forceinline__ uint32_t m_extra(uint32_t s) { return s < 16 ? 0u : (s == 16 ? 3u : (s == 17 ? 5u : (s == 18 ? 8u : 11u))); }
// L_BASE_VALUE = 0..15,16,20,28,60; M_BASE_VALUE = 0..15,16,24,56,312 (fse/constants.rs:132-134,164-166)
__device__ __forceinline__ uint32_t l_base(uint32_t s) { return s < 16 ? s : ((0x3C1C1410u >> ((s - 16) * 8)) & 0xFFu); }
__device__ __forceinline__ uint32_t m_base(uint32_t s) { return s < 16 ? s : (uint32_t)((0x0138003800180010ull >> ((s - 16) * 16)) & 0xFFFFu); }
// D_BASE_VALUE[s] = ((4 + (s & 3)) << (s >> 2)) - 4, D_EXTRA_BITS[s] = s >> 2 (fse/constants.rs:305-321): computed from the entry

template <int KIND>  // 0 = L, 1 = M, 2 = D; returns the sum of the weights (writes are clipped to the kind's states)
__device__ __forceinline__ uint32_t build_v_block(WeightReader &r, uint32_t *tab, uint32_t lane, uint32_t n_sym, uint32_t n_states,
                                                  uint32_t offset) {
    const uint32_t n_clz = __clz(n_states);
    uint32_t total = 0;
    for (uint32_t sym = 0; sym < n_sym; sym++) {
        uint32_t w = r.next();
        if (w == 0) continue;
        uint32_t k = __clz(w) - n_clz;
        uint32_t x = ((n_states << 1) >> k) - w;
        const uint32_t vb = KIND == 0 ? l_extra(sym) : (KIND == 1 ? m_extra(sym) : (sym >> 2));
        const uint32_t hi = KIND == 0 ? l_base(sym) : (KIND == 1 ? m_base(sym) : 4u + (sym & 3u));  // D: v_base = (hi << v_bits) - 4
        const uint32_t room = total < n_states ? n_states - total : 0u;
        const uint32_t wn = w < room ? w : room;
        for (uint32_t j = 0; j < wn; j++) {
            uint32_t kk, delta;
            if (j < x) { kk = k; delta = ((w + j) << k) - n_states; }
            else { kk = k - 1; delta = (j - x) << (k - 1); }
            tab[(offset + total + j) * 32 + lane] = delta | (kk << 8) | (vb << 12) | (hi << 16);
        }
        total += w;
    }
    // latch (fse/decoder.rs:288-291): k = 0, no value bits, v_base = 0 -- which for D is the field value 4: (4 << 0) - 4
    for (uint32_t t = total; t < n_states; t++) tab[(offset + t) * 32 + lane] = t | (KIND == 2 ? (4u << 16) : 0u);
    return total;
}

__global__ void __launch_bounds__(kLmdWarps * 32, 1)
k_fse_lmds(const uint8_t *__restrict__ src_base, const uint64_t *__restrict__ src_off, const uint64_t *__restrict__ src_len,
           const uint64_t *__restrict__ dst_off, const uint64_t *__restrict__ dst_cap,
           const BlockDesc *__restrict__ blocks, FseDesc *__restrict__ fse, uint32_t n_fse, LmdRec *__restrict__ lmd_scratch, uint32_t *err,
           uint32_t *work_counter) {
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
    uint32_t *tab = reinterpret_cast<uint32_t *>(smem + warp * kLmdSmemPerWarp);
    const uint32_t ring_addr = (uint32_t)__cvta_generic_to_shared(smem + warp * kLmdSmemPerWarp + 384 * 32 * 4) + lane * kRingStride;
    for (;;) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(work_counter, 32u);
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (base >= n_fse) break;
        const uint32_t f = base + lane;
        if (f < n_fse) {
            FseDesc fd = fse[f];
            const BlockDesc bd = blocks[fd.block];
            const uint8_t *blk = src_base + bd.src_off;
            const bool v1 = fd.flags & FSE_V1;
            const uint8_t *wp = blk + (v1 ? kV1HeaderSize : kV2HeaderSize);
            const uint32_t kb = bd.index < 0x1FFFFFu ? bd.index : 0x1FFFFFu;
            // The literal stage reports weight errors; this stage only needs to know the tables are sound.
            bool tables_ok = false;
            if (!(fd.flags & (FSE_TRUNC_LIT | FSE_TRUNC_LMD))) {
                // one pass over the weight payload: the L/M/D weights build their tables as they are read, the literal
                // weights are only summed; a payload that fails the reference's checks leaves the block to the literal
                // stage's error report
                WeightReader r;
                r.init(wp, fd.n_weight_bytes, v1);
                const uint32_t tl = build_v_block<0>(r, tab, lane, kLSymbols, kLStates, 0);
                const uint32_t tm = build_v_block<1>(r, tab, lane, kMSymbols, kMStates, 64);
                const uint32_t td = build_v_block<2>(r, tab, lane, kDSymbols, kDStates, 128);
                uint32_t tu = 0;
                for (uint32_t k = 0; k < 256; k++) tu += r.next();
                tables_ok = r.finish() == 0 && check_weight_totals(tl, tm, td, tu) == 0;
            }
            if (tables_ok) {

                const uint8_t *s_lo = src_base + src_off[bd.stream], *s_hi = s_lo + src_len[bd.stream];
                RingWindow br;
                int st = br.init(blk + fd.header_size + fd.n_lit_payload, fd.n_lmd_payload, fd.lmd_bits, s_lo, s_hi, ring_addr);
                if (st) {
                    atomicMin(&err[bd.stream], err_key(kb, PH_LMD, st));
                } else {
                    uint32_t sl = fd.lmd_state[0], sm = fd.lmd_state[1], sd = fd.lmd_state[2];  // relative states
                    uint32_t lit_index = 0, n_match = 0, D = 0;
                    // Stream positions in 32 bits: distances are < 2^18 and a block adds < 2^25, so clamping
                    // the bytes already produced (and the capacity left) at 2^30 changes no comparison.
                    const uint64_t pos0 = bd.dst_off - dst_off[bd.stream];
                    const uint64_t cap = dst_cap[bd.stream];
                    const uint32_t before = (uint32_t)(pos0 < 0x40000000ull ? pos0 : 0x40000000ull);
                    const uint64_t room64 = cap > pos0 ? cap - pos0 : 0;
                    const uint32_t room = (uint32_t)(room64 < 0x40000000ull ? room64 : 0x40000000ull);
                    uint32_t rel = 0;  // bytes this block has produced
                    LmdRec *out = lmd_scratch + fd.lmd_off;
                    asm volatile("" : "+l"(out));  // one register pair, not scratch base + offset at every store
                    int fail = 0;
                    const uint32_t *tl = tab + lane, *tm = tab + 64 * 32 + lane, *td = tab + 128 * 32 + lane;
                    // ---- fast path --------------------------------------------------------------------
                    // literal_index and the bytes produced only grow, so "literal_index > 40000" and "does not
                    // fit dst" are decided once at the end; per LMD only the distance is checked.  Anything
                    // suspicious (or the reader getting within 57 bits of the pad) falls back to the exact loop
                    // below, which reports what the reference reports first.
                    bool suspicious = false;
                    {
                        auto fast = [&]() -> uint2 {
                            int cur;
                            const uint64_t win = br.window_fast(cur);
                            const uint32_t el = tl[sl * 32], em = tm[sm * 32], ed = td[sd * 32];
                            const uint32_t kl = bfe(el, 8, 4), vl = bfe(el, 12, 4), km = bfe(em, 8, 4), vm = bfe(em, 12, 4);
                            const uint32_t kd_ = bfe(ed, 8, 4), vd = bfe(ed, 12, 4);
                            const int a0 = cur - (int)kl, a1 = a0 - (int)vl, a2 = a1 - (int)km, a3 = a2 - (int)vm, a4 = a3 - (int)kd_,
                                      a5 = a4 - (int)vd;
                            sl = bits_at(win, a0, kl) + (el & 0xFF);
                            const uint32_t L = (el >> 16) + bits_at(win, a1, vl);
                            sm = bits_at(win, a2, km) + (em & 0xFF);
                            const uint32_t M = (em >> 16) + bits_at(win, a3, vm);
                            sd = bits_at(win, a4, kd_) + (ed & 0xFF);
                            const uint32_t dp = ((ed >> 16) << vd) - 4u + bits_at(win, a5, vd);
                            br.P -= cur - a5;
                            D = dp ? dp : D;
                            lit_index += L;
                            rel += L;
                            suspicious |= (M != 0) & (D - 1u >= before + rel);  // D == 0 or D beyond what exists
                            rel += M;
                            return make_uint2(L | (M << 16), D);
                        };
                        uint32_t i = 0;
                        for (; i + 2 <= fd.n_lmds && br.P >= 57 + 54; i += 2) {  // two 8-byte records per 16-byte store
                            br.refill();  // two steps consume at most 13.5 bytes
                            const uint2 r0 = fast(), r1 = fast();
                            __stcg(reinterpret_cast<uint4 *>(out + i), make_uint4(r0.x, r0.y, r1.x, r1.y));
                        }
                        for (; i < fd.n_lmds && br.P >= 57; i++) { br.refill(); *reinterpret_cast<uint2 *>(out + i) = fast(); }
                        suspicious |= i != fd.n_lmds || lit_index > kLiteralsPerBlock || rel > room;
                        n_match = rel - lit_index;
                    }
                    if (suspicious) {
                        // ---- exact path: one step = one LMD (fse_core.rs:104-131), failures latched in order ----
                        br.init(blk + fd.header_size + fd.n_lit_payload, fd.n_lmd_payload, fd.lmd_bits, s_lo, s_hi, ring_addr);
                        sl = fd.lmd_state[0]; sm = fd.lmd_state[1]; sd = fd.lmd_state[2];
                        lit_index = 0; n_match = 0; D = 0; rel = 0;
                        auto step = [&]() -> uint2 {
                            int cur;
                            const uint64_t win = br.window(cur);
                            const uint32_t el = tl[sl * 32], em = tm[sm * 32], ed = td[sd * 32];
                            const uint32_t kl = bfe(el, 8, 4), vl = bfe(el, 12, 4), km = bfe(em, 8, 4), vm = bfe(em, 12, 4);
                            const uint32_t kd_ = bfe(ed, 8, 4), vd = bfe(ed, 12, 4);
                            const int a0 = cur - (int)kl, a1 = a0 - (int)vl, a2 = a1 - (int)km, a3 = a2 - (int)vm, a4 = a3 - (int)kd_,
                                      a5 = a4 - (int)vd;
                            // state bits are pulled first, then the value bits (fse/decoder.rs:214-219)
                            sl = bits_at(win, a0, kl) + (el & 0xFF);
                            const uint32_t L = (el >> 16) + bits_at(win, a1, vl);
                            sm = bits_at(win, a2, km) + (em & 0xFF);
                            const uint32_t M = (em >> 16) + bits_at(win, a3, vm);
                            sd = bits_at(win, a4, kd_) + (ed & 0xFF);
                            const uint32_t dp = ((ed >> 16) << vd) - 4u + bits_at(win, a5, vd);
                            br.P -= cur - a5;
                            D = dp ? dp : D;  // lmd/lmd_type.rs:155-159
                            lit_index += L;
                            rel += L;
                            int f1 = lit_index > kLiteralsPerBlock ? LZFSE_B200_FSE_BAD_LMD_PAYLOAD
                                     : (rel > room ? LZFSE_B200_BUFFER_OVERFLOW : 0);  // C-ABI: fixed-size Vec
                            // lz/writer.rs:156-177: distance 0 or beyond what has been written
                            const int f2 = (M != 0 && (D == 0 || D > before + rel)) ? LZFSE_B200_BAD_D_VALUE : 0;
                            n_match += M;
                            rel += M;
                            const int f3 = rel > room ? LZFSE_B200_BUFFER_OVERFLOW : 0;
                            f1 = f1 ? f1 : (f2 ? f2 : f3);
                            fail = fail ? fail : f1;
                            return make_uint2(L | (M << 16), D);
                        };
                        for (uint32_t i = 0; i < fd.n_lmds; i++) *reinterpret_cast<uint2 *>(out + i) = step();
                    }
                    if (!fail) { int cur; br.window(cur); }  // the final flush (sets nothing, P unchanged)
                    if (!fail && br.underflow()) fail = LZFSE_B200_PAYLOAD_UNDERFLOW;
                    if (!fail && !(lit_index <= fd.n_literals && n_match + lit_index == fd.n_raw && sl == 0 && sm == 0 && sd == 0))
                        fail = LZFSE_B200_FSE_BAD_LMD_PAYLOAD;
                    if (fail) atomicMin(&err[bd.stream], err_key(kb, PH_LMD, fail));
                    else fse[f].ok_lmd = 1;
                }
            }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// Expansion stage (legacy): one warp per stream walks its blocks in order.
// ------------------------------------------------------------------------------------------------
constexpr uint32_t kShortCopy = 16;  // per-lane copies up to this many bytes; longer ones go warp-wide

constexpr uint32_t kStageBytes = 512;              // a step producing at most this much is assembled in shared memory first
constexpr uint32_t kStageStride = kStageBytes + 32;  // + up to 15 bytes of alignment in front, 16-byte multiple

__device__ __forceinline__ void sts_u8(uint32_t a, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t lds_u8(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}

// One step of a bvx1/bvx2 block: 32 LMDs, one per lane (fse_core.rs:108-129, lz/writer.rs:115-180).
//
// STAGED: the step's literals and independent matches are first written into a per-warp shared-memory buffer laid out
// like the output (same address modulo 16) and then flushed with 16-byte stores.  Byte stores straight to the output
// cost an L1 sector lookup per touched sector and instruction -- after the source side was fixed they were the larger
// half of the L1 traffic that bounds this kernel.
template <bool STAGED>
__device__ __forceinline__ void expand_step(uint8_t *__restrict__ out, const uint8_t *__restrict__ lit, uint32_t L, uint32_t M, uint32_t D,
                                            uint32_t my_out, uint32_t my_lit, uint32_t out_base, uint32_t tot_o, uint32_t stage_s, uint32_t lane) {
    const uint32_t my_dst = my_out + L;  // where my match goes
    const uint32_t align = STAGED ? (uint32_t)(reinterpret_cast<uintptr_t>(out + out_base) & 15u) : 0u;
    const uint32_t sbase = stage_s + align - out_base;  // shared address of block offset 0 (only offsets of this step are used)
    // ---- match sources, requested first ----
    // A lane may copy its match on its own when everything it reads was final before this step started (or is its own
    // output); the rest go one at a time, in order, with the whole warp copying.  The independent ones' source words
    // are requested before the literals are copied so that the two round trips overlap (the kernel waits on memory).
    // The source is fetched as aligned 32-bit words (at most five cover 16 bytes at any alignment) and realigned with
    // funnel shifts: a byte load per source byte made every byte its own L1 sector lookup, and the L1 pipe, not HBM,
    // was this kernel's bound (65 % of its peak, 10 sectors per request).
    const uint8_t *src = out + my_dst - D;  // may point before `out` (earlier blocks of the stream)
    const int64_t src_rel = (int64_t)my_dst - (int64_t)D;
    const int64_t end_nonself = (src_rel + (int64_t)M < (int64_t)my_dst) ? src_rel + (int64_t)M : (int64_t)my_dst;
    const bool indep = end_nonself <= (int64_t)out_base;
    const bool solo = M != 0 && M <= kShortCopy && indep && D >= M;
    const uint32_t sm_ = solo ? M : 0u;
    const uint32_t max_m = __reduce_max_sync(0xFFFFFFFFu, sm_);
    const uint32_t src_r = (uint32_t)reinterpret_cast<uintptr_t>(src) & 3u;
    uint32_t w[5];
    {
        const uint32_t *a4 = reinterpret_cast<const uint32_t *>(reinterpret_cast<uintptr_t>(src) - src_r);
        const uint32_t nw = sm_ ? (src_r + sm_ + 3) >> 2 : 0u;
#pragma unroll
        for (uint32_t j = 0; j < 5; j++) {
            w[j] = 0;
            if (j * 4 < max_m + 6 && j < nw) w[j] = a4[j];
        }
    }
    // ---- literals ----
    const bool long_l = L > kShortCopy;
    {   // short runs: every lane copies its own; groups beyond the longest short run of the step are skipped warp-wide
        const uint32_t sl = long_l ? 0u : L;
        const uint32_t max_l = __reduce_max_sync(0xFFFFFFFFu, sl);
        const uint8_t *ps = lit + my_lit;
        uint8_t *pd = out + my_out;
        // keep the two addresses in registers: left alone, the compiler rebuilds them from the kernel parameters for
        // every predicated byte (three extra instructions per literal byte in the r1b profile)
        asm volatile("" : "+l"(ps), "+l"(pd));
#pragma unroll
        for (uint32_t g = 0; g < kShortCopy; g += 4) {
            if (g < max_l) {
                uint32_t tmp[4];
#pragma unroll
                for (uint32_t k = 0; k < 4; k++)
                    if (g + k < sl) tmp[k] = __ldg(ps + g + k);
#pragma unroll
                for (uint32_t k = 0; k < 4; k++)
                    if (g + k < sl) {
                        if (STAGED) sts_u8(sbase + my_out + g + k, tmp[k]);
                        else pd[g + k] = tmp[k];
                    }
            }
        }
    }
    uint32_t mask = __ballot_sync(0xFFFFFFFFu, long_l);
    while (mask) {
        int j = __ffs(mask) - 1;
        mask &= mask - 1;
        uint32_t o = __shfl_sync(0xFFFFFFFFu, my_out, j), li = __shfl_sync(0xFFFFFFFFu, my_lit, j), n = __shfl_sync(0xFFFFFFFFu, L, j);
        for (uint32_t t = lane; t < n; t += 32) {
            if (STAGED) sts_u8(sbase + o + t, lit[li + t]);
            else out[o + t] = lit[li + t];
        }
    }
    if (!STAGED) __syncwarp();

    // ---- matches ----
    {
        if (max_m != 0) {
            uint32_t v[4];
#pragma unroll
            for (uint32_t j = 0; j < 4; j++) v[j] = __funnelshift_r(w[j], w[j + 1], src_r * 8);
            uint8_t *pd = out + my_dst;
#pragma unroll
            for (uint32_t g = 0; g < kShortCopy; g += 4) {
                if (g < max_m) {
#pragma unroll
                    for (uint32_t k = 0; k < 4; k++)
                        if (g + k < sm_) {
                            if (STAGED) sts_u8(sbase + my_dst + g + k, v[g >> 2] >> (8 * k));
                            else pd[g + k] = (uint8_t)(v[g >> 2] >> (8 * k));
                        }
                }
            }
        }
    }
    mask = __ballot_sync(0xFFFFFFFFu, M != 0 && !solo);
    if constexpr (STAGED) {
        // ---- the other matches (long, overlapping, or reading what this step produces), in order, INTO THE IMAGE ----
        // A source byte at or above out_base is in the image (shared memory, ~30 cycles); below it, it is final and comes
        // from global memory.  Before, these matches were copied global -> global after the flush: every one of them was a
        // round trip to L2 (the line was only ever written by this SM, so it is not in L1) in the step's serial chain.
        if (mask) __syncwarp();
        while (mask) {
            const int j = __ffs(mask) - 1;
            mask &= mask - 1;
            const uint32_t o = __shfl_sync(0xFFFFFFFFu, my_dst, j), d = __shfl_sync(0xFFFFFFFFu, D, j), n = __shfl_sync(0xFFFFFFFFu, M, j);
            const uint8_t *s = out + o - d;   // may point before `out` (earlier blocks of the stream)
            const int32_t src0 = (int32_t)o - (int32_t)d, ob = (int32_t)out_base;  // block-relative; src0 may be negative
            for (uint32_t t = lane; t < n; t += 32) {
                const uint32_t u = d >= n ? t : t % d;  // byte i == byte i mod D of the D bytes before the match
                const int32_t pp = src0 + (int32_t)u;
                const uint32_t v = pp >= ob ? lds_u8(sbase + (uint32_t)pp) : (uint32_t)s[u];
                sts_u8(sbase + o + t, v);
            }
            __syncwarp();
        }
        // flush [out_base, out_base + tot_o): bytes up to the first 16-byte boundary, whole 16-byte units, the rest
        __syncwarp();
        uint8_t *g0 = out + out_base;
        asm volatile("" : "+l"(g0));  // as above: one address, not one per store
        const uint32_t end = align + tot_o;
        const uint32_t body_lo = (align + 15) & ~15u, body_hi = end & ~15u;
        if (body_lo <= body_hi) {
            if (lane < body_lo - align) g0[lane] = (uint8_t)lds_u8(stage_s + align + lane);
            for (uint32_t c = body_lo + lane * 16; c < body_hi; c += 512) {
                uint4 q;
                asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "r"(stage_s + c) : "memory");
                *reinterpret_cast<uint4 *>(g0 + (c - align)) = q;
            }
            if (lane < end - body_hi) g0[body_hi - align + lane] = (uint8_t)lds_u8(stage_s + body_hi + lane);
        } else if (lane < tot_o) {
            g0[lane] = (uint8_t)lds_u8(stage_s + align + lane);
        }
        __syncwarp();
    } else {
        if (mask) __syncwarp();
        while (mask) {
            int j = __ffs(mask) - 1;
            mask &= mask - 1;
            uint32_t o = __shfl_sync(0xFFFFFFFFu, my_dst, j), d = __shfl_sync(0xFFFFFFFFu, D, j), n = __shfl_sync(0xFFFFFFFFu, M, j);
            const uint8_t *s = out + o - d;
            if (d >= n) {
                for (uint32_t t = lane; t < n; t += 32) out[o + t] = s[t];
            } else {
                for (uint32_t t = lane; t < n; t += 32) out[o + t] = s[t % d];  // byte i == byte i mod D of the seed
            }
            __syncwarp();
        }
        __syncwarp();
    }
}

// One bvx1/bvx2 block, 32 LMDs per step.
__device__ void expand_fse_block(uint8_t *__restrict__ out /* block's first output byte */, const uint8_t *__restrict__ lit,
                                 const LmdRec *__restrict__ lmds, uint32_t n_lmds, uint32_t stage_s, uint32_t lane) {
    uint32_t out_base = 0, lit_base = 0;  // running offsets inside the block
    asm volatile("" : "+l"(out), "+l"(lit), "+l"(lmds));  // plain register pointers (see expand_step)
    // The records of step b+1 are fetched while step b is being copied: one memory round trip less per step.
    uint2 nxt = make_uint2(0, 0);
    if (lane < n_lmds) nxt = __ldg(reinterpret_cast<const uint2 *>(lmds) + lane);
    for (uint32_t b = 0; b < n_lmds; b += 32) {
        const uint2 rec = nxt;
        nxt = make_uint2(0, 0);
        if (b + 32 + lane < n_lmds) nxt = __ldg(reinterpret_cast<const uint2 *>(lmds) + b + 32 + lane);
        const uint32_t L = rec.x & 0xFFFF, M = rec.x >> 16, D = rec.y;
        if (lane == 0) asm volatile("prefetch.global.L1 [%0];" ::"l"(lit + lit_base + 128));  // scratch has slack past its end
        // inclusive scan of (sum L) << 17 | (sum L+M): 32*315 < 2^14, 32*(315+2359) < 2^17
        uint32_t v = (L << 17) + (L + M), inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
            if (lane >= (uint32_t)o) inc += t;
        }
        const uint32_t exc = inc - v;
        const uint32_t tot = __shfl_sync(0xFFFFFFFFu, inc, 31);
        const uint32_t my_out = out_base + (exc & 0x1FFFF);  // where my literals go
        const uint32_t my_lit = lit_base + (exc >> 17);
        const uint32_t tot_o = tot & 0x1FFFF;
        if (tot_o <= kStageBytes) expand_step<true>(out, lit, L, M, D, my_out, my_lit, out_base, tot_o, stage_s, lane);
        else expand_step<false>(out, lit, L, M, D, my_out, my_lit, out_base, tot_o, stage_s, lane);
        out_base += tot_o;
        lit_base += tot >> 17;
    }
}

#ifndef LZB_EXPAND_CTAS
#define LZB_EXPAND_CTAS 4   // resident 8-warp CTAs per SM the register budget is set for
#endif
constexpr int kExpandWarps = 8;

__global__ void __launch_bounds__(kExpandWarps * 32, LZB_EXPAND_CTAS)
k_expand(const uint8_t *__restrict__ src_base, const uint64_t *__restrict__ src_off, const uint64_t *__restrict__ src_len,
         uint8_t *__restrict__ dst_base, const uint64_t *__restrict__ dst_off, const uint64_t *__restrict__ dst_cap,
         const StreamCounts *__restrict__ bases,
         const BlockDesc *__restrict__ blocks, const FseDesc *__restrict__ fse, const uint8_t *__restrict__ lit_scratch,
         const LmdRec *__restrict__ lmd_scratch, uint32_t *err, size_t n_streams, uint32_t *work_counter, const uint64_t *__restrict__ long_base) {
    __shared__ __align__(16) uint8_t stage[kExpandWarps][kStageStride];
    const uint32_t lane = lane_id();
    uint32_t stage_s = (uint32_t)__cvta_generic_to_shared(stage[threadIdx.x >> 5]);
    asm volatile("" : "+r"(stage_s));  // or the compiler re-derives it from %tid and the shared window base in every step
  // Persistent warps pull streams from a counter: with a fixed stream per warp a CTA's eight slots stay occupied until
  // its slowest stream is done.
  for (;;) {
    size_t stream = 0;
    if (lane == 0) stream = atomicAdd(work_counter, 1u);
    stream = __shfl_sync(0xFFFFFFFFu, (uint32_t)stream, 0);
    if (stream >= n_streams) return;
    if (long_base[stream] != ~0ull) continue;  // expand_long.cu
    const uint64_t b0 = bases[stream].n_blocks, b1 = bases[stream + 1].n_blocks;
    uint8_t *stream_out = dst_base + dst_off[stream];
    if (b1 - b0 == 1 && vn_fast_eligible(1, blocks[b0], src_off[stream] + src_len[stream] - blocks[b0].src_off, dst_cap[stream])) continue;  // k_expand_vn
    for (uint64_t b = b0; b < b1; b++) {
        const BlockDesc bd = blocks[b];
        uint8_t *out = dst_base + bd.dst_off;
        const uint32_t kb = bd.index < 0x1FFFFFu ? bd.index : 0x1FFFFFu;
        if (bd.type == BT_RAW) {
            warp_copy(out, src_base + bd.src_off + 8, bd.n_raw, lane);
        } else if (bd.type == BT_VXN) {
            int st = 0;
            if (lane == 0) {
                const uint64_t rest = src_off[stream] + src_len[stream] - bd.src_off;
                const uint64_t pos = bd.dst_off - dst_off[stream];
                st = vn_decode_block(src_base + bd.src_off, rest, stream_out, pos, dst_cap[stream]);
                if (st) atomicMin(&err[stream], err_key(kb, PH_LMD, st));
            }
            st = __shfl_sync(0xFFFFFFFFu, st, 0);
            if (st) break;
        } else {
            const FseDesc &fd = fse[bd.fse_idx];
            if (!(fd.ok_lit && fd.ok_lmd)) break;  // the entropy stages already recorded why
            expand_fse_block(out, lit_scratch + fd.lit_off, lmd_scratch + fd.lmd_off, fd.n_lmds, stage_s, lane);
        }
        __syncwarp();
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------
// Single-LZVN-block streams (vn_fast_eligible): a warp per stream, 32 payload bytes per step.
//
// The opcode stream is a serial chain -- an opcode's position depends on the length of the one before -- but the
// chain is cheap to resolve in parallel: every lane decodes "the opcode that would start at payload byte p + lane"
// (class, lengths, distance: one table lookup and a few selects), the lanes that really are opcode starts are found
// by pointer doubling over next = lane + opcode length (five rounds), distances inherited from the previous opcode
// come from a ballot, output positions from a warp scan.  The real opcodes of the window (about eight on text) then
// copy side by side: literals from the payload to a shared-memory image of the output, matches inside the image
// (independent short ones per lane; long, overlapping or in-window-dependent ones in order, warp-wide), and the
// finished image goes to HBM once in 16-byte stores.
// History (512 MiB of 21..4096-byte inputs): one lane interpreting the block against HBM, what the in-order kernels
// do, 20.8 ms; a warp per stream with every lane decoding the same opcode 11.9 ms (bound by instruction issue, ~100
// per opcode); eight lanes per stream, four streams per warp 15.8 ms (the groups diverge on the opcode class).
// Anything irregular (bad opcode, short payload, distance before the start, counts that do not add up) abandons the
// image and re-runs the exact interpreter, which reports what the reference reports (vn/vn_core.rs:51-140).
// ------------------------------------------------------------------------------------------------
constexpr int kVnWarps = 8;
constexpr uint32_t kVnImage = kVnFastRaw + 16;  // + up to 15 bytes in front so that image and output agree modulo 16
constexpr uint32_t kVnSolo = 16;                // per-lane copies up to this many bytes

// Opcode table entry: oplen[0:2] | inline L[2:6] | inline M[6:12] | L += byte1 + 16 [12] | M += byte1 + 16 [13] |
// distance kind [14:16] (0 none, 1 SmlD, 2 MedD, 3 LrgD) | Eos [16] | undefined [17]   (vn/opc.rs:18-228)
__device__ uint32_t vn_table_entry(uint32_t b) {
    switch (vn_op(b)) {
    case OP_SML_L: return 1u | ((b & 0xF) << 2);
    case OP_LRG_L: return 2u | (1u << 12);
    case OP_SML_M: return 1u | ((b & 0xF) << 6);
    case OP_LRG_M: return 2u | (1u << 13);
    case OP_PRE_D: return 1u | (((b >> 6) & 3) << 2) | ((((b >> 3) & 7) + 3) << 6);
    case OP_SML_D: return 2u | (((b >> 6) & 3) << 2) | ((((b >> 3) & 7) + 3) << 6) | (1u << 14);
    case OP_MED_D: return 3u | (((b >> 3) & 3) << 2) | ((((b & 7) << 2) + 3) << 6) | (2u << 14);
    case OP_LRG_D: return 3u | (((b >> 6) & 3) << 2) | ((((b >> 3) & 7) + 3) << 6) | (3u << 14);
    case OP_NOP: return 1u;
    case OP_EOS: return 1u << 16;
    default: return 1u << 17;
    }
}
__device__ __forceinline__ uint32_t lanemask_le() { uint32_t m; asm("mov.u32 %0, %%lanemask_le;" : "=r"(m)); return m; }

#ifndef LZB_VN_CTAS
#define LZB_VN_CTAS 4
#endif
__global__ void __launch_bounds__(kVnWarps * 32, LZB_VN_CTAS)
k_expand_vn(const uint8_t *__restrict__ src_base, const uint64_t *__restrict__ src_off, const uint64_t *__restrict__ src_len,
            uint8_t *__restrict__ dst_base, const uint64_t *__restrict__ dst_off, const uint64_t *__restrict__ dst_cap,
            const StreamCounts *__restrict__ bases, const BlockDesc *__restrict__ blocks, uint32_t *err, size_t n_streams,
            uint32_t *work_counter) {
    __shared__ __align__(16) uint8_t image[kVnWarps][kVnImage];
    __shared__ uint32_t optab[256];
    static_assert(kVnWarps * 32 == 256, "one table entry per thread");
    optab[threadIdx.x] = vn_table_entry(threadIdx.x);
    __syncthreads();
    constexpr uint32_t kFull = 0xFFFFFFFFu;
    const uint32_t lane = lane_id();
  // Persistent warps pull streams from a counter: inputs of 21..4096 bytes take very different times, and a CTA that
  // waits for its longest stream left a third of the resident warps idle.
  for (;;) {
    size_t stream = 0;
    if (lane == 0) stream = atomicAdd(work_counter, 1u);
    stream = __shfl_sync(kFull, (uint32_t)stream, 0);
    if (stream >= n_streams) return;
    const uint64_t b0 = bases[stream].n_blocks, nb = bases[stream + 1].n_blocks - b0;
    if (nb != 1) continue;
    const BlockDesc bd = blocks[b0];
    const uint64_t src_rest = src_off[stream] + src_len[stream] - bd.src_off, cap = dst_cap[stream];
    if (!vn_fast_eligible(nb, bd, src_rest, cap)) continue;
    const uint8_t *src = src_base + bd.src_off;
    uint8_t *out = dst_base + dst_off[stream];
    const uint32_t n_raw = bd.n_raw, n_payload = ld_u32(src + 8);
    const uint32_t vlen = (uint32_t)src_rest - kVnHeaderSize;
    const uint32_t bias = (uint32_t)(reinterpret_cast<uintptr_t>(out) & 15u);
    const uint32_t img = (uint32_t)__cvta_generic_to_shared(image[threadIdx.x >> 5]) + bias;  // shared address of output byte 0
    uint32_t used = 0, out_pos = 0, D = 0;
    bool ok = true, eos = false;
    while (ok && !eos) {
        // ---- what would an opcode starting at payload byte used + lane be? ----
        const uint8_t *s = src + kVnHeaderSize + used + lane;
        const uint32_t rem = vlen - used - lane;  // wraps for lanes past the end: caught by `alive`
        const bool alive = used + lane + 8 <= vlen;  // an opcode may only start where at least 8 bytes are left (vn_core.rs:155-160)
        const uint32_t w = alive ? ldg4u(s) : 0u;
        const uint32_t e = optab[w & 0xFF], b1 = (w >> 8) & 0xFF, dk = (e >> 14) & 3;
        const uint32_t oplen = e & 3;
        const uint32_t L = ((e >> 2) & 15) + ((e >> 12) & 1) * (b1 + 16);
        const uint32_t M = ((e >> 6) & 63) + ((e >> 13) & 1) * (b1 + 16) + (dk == 2 ? (b1 & 3) : 0u);
        const uint32_t d_new = dk == 1 ? (((w & 7) << 8) | b1) : (dk == 2 ? ((w >> 10) & 0x3FFF) : ((w >> 8) & 0xFFFF));
        const bool is_eos = alive && ((e >> 16) & 1);
        const bool bad = !alive || ((e >> 17) & 1) || (!is_eos && rem - oplen < L + 8);
        // ---- which lanes are opcode starts: the orbit of lane 0 under next = lane + length ----
        uint32_t jump = (bad || is_eos) ? 64u : lane + oplen + L;  // the chain stops at an Eos or a bad opcode
        const uint32_t next = jump;
        uint32_t reach = 1u;
#pragma unroll
        for (int r = 0; r < 5; r++) {
            const uint32_t c = (((reach >> lane) & 1u) && jump < 32u) ? (1u << jump) : 0u;
            reach |= __reduce_or_sync(kFull, c);
            const uint32_t j2 = __shfl_sync(kFull, jump, jump & 31u);
            jump = jump < 32u ? j2 : jump;
        }
        const bool real = (reach >> lane) & 1u;
        if (__any_sync(kFull, real && bad)) { ok = false; break; }
        const uint32_t real_m = reach;
        const int last = 31 - __clz(real_m);                       // the window's last opcode
        const uint32_t eos_m = __ballot_sync(kFull, real && is_eos);  // at most the last one
        // ---- distances: an opcode without one inherits the previous opcode's ----
        const uint32_t d_m = __ballot_sync(kFull, real && dk != 0);
        const uint32_t below = d_m & lanemask_le();
        const uint32_t d_src = __shfl_sync(kFull, d_new, below ? 31 - __clz(below) : 0);
        const uint32_t my_d = below ? d_src : D;
        // ---- output positions ----
        const uint32_t n_out = (real && !is_eos) ? L + M : 0u;
        uint32_t inc = n_out;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(kFull, inc, o);
            if (lane >= (uint32_t)o) inc += t;
        }
        const uint32_t total = __shfl_sync(kFull, inc, 31);
        const uint32_t my_out = out_pos + inc - n_out, my_dst = my_out + L;
        const bool has_l = real && !is_eos && L != 0, has_m = real && !is_eos && M != 0;
        if (out_pos + total > n_raw || __any_sync(kFull, has_m && (my_d == 0 || my_d > my_dst))) { ok = false; break; }
        // ---- literals: payload -> image ----
        {
            const uint8_t *ls = s + oplen;
            const uint32_t sl = (has_l && L <= kVnSolo) ? L : 0u;
            const uint32_t max_l = __reduce_max_sync(kFull, sl);
            for (uint32_t g = 0; g < max_l; g += 4) {
                uint32_t t[4];
#pragma unroll
                for (uint32_t k = 0; k < 4; k++)
                    if (g + k < sl) t[k] = __ldg(ls + g + k);
#pragma unroll
                for (uint32_t k = 0; k < 4; k++)
                    if (g + k < sl) sts_u8(img + my_out + g + k, t[k]);
            }
            uint32_t lm = __ballot_sync(kFull, has_l && L > kVnSolo);
            while (lm) {
                const int j = __ffs(lm) - 1;
                lm &= lm - 1;
                const uint32_t o = __shfl_sync(kFull, my_out, j), n = __shfl_sync(kFull, L, j), ol = __shfl_sync(kFull, oplen, j);
                const uint8_t *p = src + kVnHeaderSize + used + j + ol;
                for (uint32_t t = lane; t < n; t += 32) sts_u8(img + o + t, __ldg(p + t));
            }
        }
        __syncwarp();
        // ---- matches: image -> image (lz/writer.rs:144-180) ----
        {
            const uint32_t from = my_dst - my_d;
            const uint32_t end_nonself = from + M < my_dst ? from + M : my_dst;
            const bool solo = has_m && M <= kVnSolo && my_d >= M && end_nonself <= out_pos;  // everything it reads was final before this window
            const uint32_t sm = solo ? M : 0u;
            const uint32_t max_m = __reduce_max_sync(kFull, sm);
            for (uint32_t g = 0; g < max_m; g += 4) {
                uint32_t t[4];
#pragma unroll
                for (uint32_t k = 0; k < 4; k++)
                    if (g + k < sm) t[k] = lds_u8(img + from + g + k);
#pragma unroll
                for (uint32_t k = 0; k < 4; k++)
                    if (g + k < sm) sts_u8(img + my_dst + g + k, t[k]);
            }
            uint32_t mm = __ballot_sync(kFull, has_m && !solo);
            if (mm) __syncwarp();
            while (mm) {  // in order, the whole warp on one match
                const int j = __ffs(mm) - 1;
                mm &= mm - 1;
                const uint32_t o = __shfl_sync(kFull, my_dst, j), d = __shfl_sync(kFull, my_d, j), n = __shfl_sync(kFull, M, j);
                if (d >= n) for (uint32_t t = lane; t < n; t += 32) sts_u8(img + o + t, lds_u8(img + o - d + t));
                else for (uint32_t t = lane; t < n; t += 32) sts_u8(img + o + t, lds_u8(img + o - d + t % d));
                __syncwarp();
            }
        }
        __syncwarp();
        // ---- advance ----
        out_pos += total;
        if (d_m) D = __shfl_sync(kFull, d_new, 31 - __clz(d_m));
        if (eos_m) {  // Eos must be 06 00 00 00 00 00 00 00 (vn_core.rs:179-187)
            const int j = __ffs(eos_m) - 1;
            const uint8_t *p = src + kVnHeaderSize + used + j;
            if (ldg4u(p) != 0x06u || ldg4u(p + 4) != 0u) { ok = false; break; }
            used += j + 8;
            eos = true;
        } else {
            used += __shfl_sync(kFull, next, last);
        }
    }
    if (ok && eos && used == n_payload && out_pos == n_raw) {  // VnCore::decode's final accounting (vn_core.rs:96-111)
        __syncwarp();
        // head up to the first 16-byte boundary of the output, 16-byte units, tail
        uint32_t head = (16u - bias) & 15u;
        if (head > n_raw) head = n_raw;
        if (lane < head) out[lane] = (uint8_t)lds_u8(img + lane);
        const uint32_t nv = (n_raw - head) >> 4;
        for (uint32_t i = lane; i < nv; i += 32) {
            uint4 q;
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "r"(img + head + i * 16) : "memory");
            *reinterpret_cast<uint4 *>(out + head + i * 16) = q;
        }
        const uint32_t done = head + nv * 16;
        if (done + lane < n_raw) out[done + lane] = (uint8_t)lds_u8(img + done + lane);
    } else if (lane == 0) {
        const int st = vn_decode_block(src, src_rest, out, 0, cap);
        const uint32_t kb = bd.index < 0x1FFFFFu ? bd.index : 0x1FFFFFu;
        if (st) atomicMin(&err[stream], err_key(kb, PH_LMD, st));
    }
    __syncwarp();  // the image is reused by the next stream
  }
}

__global__ void k_finish(const uint32_t *__restrict__ err, const uint64_t *__restrict__ raw_total, uint64_t *out_len, int32_t *status, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t k = err[i];
    int32_t st = k == kNoError ? 0 : (int32_t)(k & 0xFF);
    status[i] = st;
    if (out_len) out_len[i] = st == 0 ? raw_total[i] : 0;
}

// ------------------------------------------------------------------------------------------------
// Host-side launchers (called from api.cu)
// ------------------------------------------------------------------------------------------------
// Streams with at least this many bvx blocks take the two-pass expansion (expand_long.cu): a handful in a batch that
// cannot fill the machine with one warp per stream, many when it can (the image costs four bytes per output byte).
uint32_t long_stream_threshold(size_t n, int n_sms) {
    static const int forced = [] { const char *e = getenv("LZB_LONG_FSE"); return e ? atoi(e) : 0; }();  // measurements / tests
    if (forced > 0) return (uint32_t)forced;
    return n < (size_t)n_sms * 16 ? 4u : 64u;
}
void launch_scan_count(const uint8_t *src, const uint64_t *src_off, const uint64_t *src_len, const uint64_t *dst_cap, size_t n,
                       StreamCounts *counts, uint32_t *err, uint64_t *raw_total, uint32_t *n_blocks_out, StreamCounts *totals, uint32_t *work,
                       uint32_t long_fse, const uint64_t *limit, uint8_t *more_out, cudaStream_t s) {
    if (n == 0) return;
    const int tb = 128;
    k_scan<false><<<(unsigned)((n + tb - 1) / tb), tb, 0, s>>>(src, src_off, src_len, nullptr, dst_cap, n, counts, nullptr, nullptr, err, raw_total, n_blocks_out,
                                                               work, long_fse, nullptr, nullptr, nullptr, limit, more_out);
    launch_exclusive_scan(counts, n, totals, work, s);
}
void launch_scan_fill(const uint8_t *src, const uint64_t *src_off, const uint64_t *src_len, const uint64_t *dst_off, const uint64_t *dst_cap, size_t n,
                      StreamCounts *bases, BlockDesc *blocks, FseDesc *fse, uint32_t *err, uint64_t *raw_total, uint32_t *work, uint32_t long_fse,
                      uint64_t *long_base, uint32_t *long_blocks, uint32_t *long_streams, const uint64_t *limit, cudaStream_t s) {
    if (n == 0) return;
    const int tb = 128;
    k_scan<true><<<(unsigned)((n + tb - 1) / tb), tb, 0, s>>>(src, src_off, src_len, dst_off, dst_cap, n, bases, blocks, fse, err, raw_total, nullptr, work,
                                                              long_fse, long_base, long_blocks, long_streams, limit, nullptr);
}

// ------------------------------------------------------------------------------------------------
// Bounded decode (lzfse_b200_decode_prefix_batch_*): the blocks that cover the limit are decoded into an internal buffer
// laid out by an exclusive scan of their sizes; the prefix each caller asked for is copied out afterwards.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_scan_u64(const uint64_t *__restrict__ in, uint64_t *__restrict__ out, size_t n, uint64_t *total) {
    // one CTA: thread t owns a contiguous run, the 1024 run sums are scanned in shared memory (16-byte aligned starts)
    __shared__ uint64_t sums[1024];
    const size_t per = (n + 1023) / 1024, lo = (size_t)threadIdx.x * per, hi = lo + per < n ? lo + per : n;
    uint64_t acc = 0;
    for (size_t i = lo; i < hi; i++) acc += (in[i] + 15) & ~15ull;
    sums[threadIdx.x] = acc;
    __syncthreads();
    for (uint32_t o = 1; o < 1024; o <<= 1) {
        const uint64_t t = threadIdx.x >= o ? sums[threadIdx.x - o] : 0;
        __syncthreads();
        sums[threadIdx.x] += t;
        __syncthreads();
    }
    uint64_t run = sums[threadIdx.x] - acc;
    for (size_t i = lo; i < hi; i++) { out[i] = run; run += (in[i] + 15) & ~15ull; }
    if (threadIdx.x == 1023) *total = sums[1023];
}
__global__ void __launch_bounds__(256) k_prefix_copy(const uint8_t *__restrict__ inner, const uint64_t *__restrict__ inner_off,
                                                     const uint64_t *__restrict__ inner_len, const int32_t *__restrict__ status,
                                                     const uint64_t *__restrict__ limit, uint8_t *__restrict__ dst, const uint64_t *__restrict__ dst_off,
                                                     uint64_t *out_len, uint8_t *more, size_t n) {
    for (size_t i = blockIdx.x; i < n; i += gridDim.x) {  // a CTA per stream
        const uint64_t have = status[i] == 0 ? inner_len[i] : 0, want = limit[i];
        const uint64_t m = have < want ? have : want;
        const uint8_t *s = inner + inner_off[i];  // 16-byte aligned
        uint8_t *d = dst + dst_off[i];
        if ((reinterpret_cast<uintptr_t>(d) & 15) == 0) {
            const uint64_t nv = m >> 4;
            for (uint64_t k = threadIdx.x; k < nv; k += 256) reinterpret_cast<uint4 *>(d)[k] = reinterpret_cast<const uint4 *>(s)[k];
            for (uint64_t k = (nv << 4) + threadIdx.x; k < m; k += 256) d[k] = s[k];
        } else {
            for (uint64_t k = threadIdx.x; k < m; k += 256) d[k] = s[k];
        }
        if (threadIdx.x == 0) {
            out_len[i] = m;
            if (status[i] != 0) more[i] = 0;
            else if (have > want) more[i] = 1;  // (otherwise as the scan left it: blocks remain, or the frame ended)
        }
    }
}
void launch_scan_u64(const uint64_t *in, uint64_t *out, size_t n, uint64_t *total, cudaStream_t s) { k_scan_u64<<<1, 1024, 0, s>>>(in, out, n, total); }
void launch_prefix_copy(const uint8_t *inner, const uint64_t *inner_off, const uint64_t *inner_len, const int32_t *status, const uint64_t *limit, uint8_t *dst,
                        const uint64_t *dst_off, uint64_t *out_len, uint8_t *more, size_t n, int n_sms, cudaStream_t s) {
    if (n == 0) return;
    const size_t g = n < (size_t)n_sms * 8 ? n : (size_t)n_sms * 8;
    k_prefix_copy<<<(unsigned)g, 256, 0, s>>>(inner, inner_off, inner_len, status, limit, dst, dst_off, out_len, more, n);
}
int setup_decode_kernels() {
    cudaError_t e = cudaFuncSetAttribute(k_fse_literals, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kLitWarps * kLitSmemPerWarp));
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(k_fse_lmds, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kLmdWarps * kLmdSmemPerWarp));
    return (int)e;
}
void launch_fse_stages(const uint8_t *src, const uint64_t *src_off, const uint64_t *src_len, const uint64_t *dst_off, const uint64_t *dst_cap,
                       const BlockDesc *blocks, FseDesc *fse, uint32_t n_fse, uint8_t *lit_scratch, LmdRec *lmd_scratch, uint32_t *err,
                       uint32_t *work_counters /* kWorkWords zeroed u32 */, int n_sms, cudaStream_t s, cudaEvent_t between,
                       cudaStream_t side /* or null */, cudaEvent_t fork, cudaEvent_t join) {
    if (n_fse == 0) return;
    unsigned need_lit = (n_fse + 32 * kLitWarps - 1) / (32 * kLitWarps), need_lmd = (n_fse + 32 * kLmdWarps - 1) / (32 * kLmdWarps);
    unsigned g_lit = need_lit < (unsigned)n_sms ? need_lit : (unsigned)n_sms;
    unsigned g_lmd = need_lmd < (unsigned)n_sms ? need_lmd : (unsigned)n_sms;
    // The two stages are independent of each other, and each of their CTAs fills an SM's shared memory.  When both grids
    // fit the machine side by side (batches of few blocks, e.g. 8 streams of 16 MiB) the literal stage runs on a side stream.
    const bool side_by_side = side != nullptr && g_lit + g_lmd <= (unsigned)n_sms;
    if (side_by_side) {
        cudaEventRecord(fork, s);
        cudaStreamWaitEvent(side, fork, 0);
        k_fse_literals<<<g_lit, kLitWarps * 32, kLitWarps * kLitSmemPerWarp, side>>>(src, src_off, src_len, blocks, fse, n_fse, lit_scratch, err, work_counters);
        cudaEventRecord(join, side);
    } else {
        k_fse_literals<<<g_lit, kLitWarps * 32, kLitWarps * kLitSmemPerWarp, s>>>(src, src_off, src_len, blocks, fse, n_fse, lit_scratch, err, work_counters);
    }
    if (between) cudaEventRecord(between, s);
    k_fse_lmds<<<g_lmd, kLmdWarps * 32, kLmdWarps * kLmdSmemPerWarp, s>>>(src, src_off, src_len, dst_off, dst_cap, blocks, fse, n_fse, lmd_scratch, err, work_counters + 1);
    if (side_by_side) cudaStreamWaitEvent(s, join, 0);
}
void launch_expand(const uint8_t *src, const uint64_t *src_off, const uint64_t *src_len, uint8_t *dst, const uint64_t *dst_off,
                   const uint64_t *dst_cap, const StreamCounts *bases, const BlockDesc *blocks, const FseDesc *fse, const uint8_t *lit_scratch,
                   const LmdRec *lmd_scratch, uint32_t *err, size_t n, uint32_t *work_counter /* zeroed */, const uint64_t *long_base, int n_sms,
                   cudaStream_t s) {
    if (n == 0) return;
    // Tuning aid: LZB_EXPAND_PAD_KB adds unused dynamic shared memory per CTA and so caps the resident CTAs per SM.
    static const int pad_kb = [] { const char *e = getenv("LZB_EXPAND_PAD_KB"); return e ? atoi(e) : 0; }();
    if (pad_kb > 48) cudaFuncSetAttribute(k_expand, cudaFuncAttributeMaxDynamicSharedMemorySize, pad_kb * 1024);
    const size_t need = (n + kExpandWarps - 1) / kExpandWarps, resident = (size_t)n_sms * LZB_EXPAND_CTAS;
    k_expand<<<(unsigned)(need < resident ? need : resident), kExpandWarps * 32, (size_t)pad_kb * 1024, s>>>(src, src_off, src_len, dst, dst_off, dst_cap, bases, blocks, fse,
                                                                                                        lit_scratch, lmd_scratch, err, n, work_counter, long_base);
}
void launch_expand_vn(const uint8_t *src, const uint64_t *src_off, const uint64_t *src_len, uint8_t *dst, const uint64_t *dst_off,
                      const uint64_t *dst_cap, const StreamCounts *bases, const BlockDesc *blocks, uint32_t *err, size_t n,
                      uint32_t *work_counter /* zeroed */, int n_sms, cudaStream_t s) {
    if (n == 0) return;
    const size_t need = (n + kVnWarps - 1) / kVnWarps, resident = (size_t)n_sms * LZB_VN_CTAS;
    k_expand_vn<<<(unsigned)(need < resident ? need : resident), kVnWarps * 32, 0, s>>>(src, src_off, src_len, dst, dst_off, dst_cap, bases, blocks, err, n,
                                                                                           work_counter);
}
void launch_finish(const uint32_t *err, const uint64_t *raw_total, uint64_t *out_len, int32_t *status, size_t n, cudaStream_t s) {
    if (n == 0) return;
    const int tb = 256;
    k_finish<<<(unsigned)((n + tb - 1) / tb), tb, 0, s>>>(err, raw_total, out_len, status, n);
}

}  // namespace lzb
