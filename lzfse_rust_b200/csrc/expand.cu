// expand.cu -- LZ expansion stage of the batched decoder: one CTA per stream, output window in shared memory.
//
// What the reference does per LMD in FseCore::decode_internal (fse/fse_core.rs:108-129: copy L literals, then
// write_match(M, D) with LZ77 overlap semantics, lz/writer.rs:144-180) is done here 32 LMDs at a time per warp,
// with the last 64 KiB of the stream's output held in a shared-memory ring:
//
//   * match sources are shared-memory reads (an LZFSE match on text is ~6 bytes; fetching a 32-byte sector from
//     HBM for each of them was 5x the algorithmic traffic and all of the latency of the previous design);
//   * the ring is written to HBM exactly once, in 16-byte coalesced stores, by a dedicated flusher warp;
//   * the 7 worker warps of a CTA take consecutive 32-LMD groups of the SAME stream round-robin.  Two short serial
//     chains link them, both through shared memory: the running (output, literal) offsets (a chained scan: each
//     group adds its total to its predecessor's and publishes it), and an in-order commit watermark `pos_done`
//     ("every output byte below is final").  A match copies as soon as its source lies below the watermark; the
//     few whose source is younger wait for their group's turn and are then resolved inside the warp in rounds.
//
// Ring discipline (W = ring bytes, S = how far ahead of the flusher anyone may write, 2S + max M <= W):
//   a group may write [gb, ge) once ge <= flushed + S, so nobody overwrites a byte that is not in HBM yet, and while
//   a group based at gb is uncommitted every position >= gb + S - W is still intact in the ring; sources below that
//   bound (or below `ring_lo`, the end of the last raw/LZVN block, which bypass the ring) are read back from HBM
//   with L1 bypassed -- they were flushed because flushed >= gb - S.
#include <cstddef>
#include <type_traits>

#include "common.cuh"
#include "lz_blocks.cuh"

namespace lzb {

#ifndef LZB_XRING_LOG2
#define LZB_XRING_LOG2 16
#endif
#ifndef LZB_XWORKERS
#define LZB_XWORKERS 7
#endif
#ifndef LZB_XCTAS
#define LZB_XCTAS 3
#endif
constexpr uint32_t kXRing = 1u << LZB_XRING_LOG2, kXMask = kXRing - 1;
constexpr uint32_t kXAhead = kXRing / 4;  // S
constexpr uint32_t kXBig = 1u << 12;      // a group producing more than this is committed in sub-ranges
constexpr uint32_t kXSolo = 32;           // per-lane copies up to this many bytes; longer ones go warp-wide
constexpr int kXBatch = 8;                // bytes a lane loads before it stores them
constexpr uint32_t kXFlushMin = 512;
constexpr int kXWorkers = LZB_XWORKERS;
constexpr int kXSlots = 16;               // carry slots, >= 2 * workers
constexpr int kXThreads = (kXWorkers + 1) * 32;
constexpr uint32_t kFull = 0xFFFFFFFFu;
static_assert(2 * kXAhead + 4096 <= kXRing, "ring discipline");
static_assert(kXBig <= kXAhead / 2, "the oldest uncommitted group must always be allowed to write");
static_assert(kXBatch == 8, "solo_copy is written out for 8");
static_assert(kXSlots >= 2 * kXWorkers && (kXSlots & (kXSlots - 1)) == 0, "carry slots");

struct XCtl {
    unsigned long long carry[kXSlots];  // (out_end << 32) | (lit_end << 16) | tag of the group that wrote it
    uint32_t pos_done;                  // every stream byte below this position is final (in the ring or in HBM)
    uint32_t done_g;                    // number of groups this CTA has committed (numbering runs on across streams)
    uint32_t flushed;                   // every stream byte below this position is in HBM
    uint32_t ring_lo;                   // positions below were produced by raw/LZVN blocks (HBM only)
    uint32_t stream;
    int32_t stop;
};
constexpr size_t kXSmem = kXRing + sizeof(XCtl);
#define XOFF(field) ((uint32_t)offsetof(XCtl, field))

extern __shared__ __align__(16) uint8_t x_smem[];

// Everything below addresses shared memory through 32-bit shared-window addresses and inline PTX: with generic
// pointers (or the array symbol under a predicate) the compiler re-derives the window base for every byte, which
// made the copy loops 3.5x longer than they have to be.
//
// Flags and data both live in shared memory, whose accesses one SM performs in issue order, so the workers publish
// with plain (volatile) stores after a __syncwarp and the waiters poll with volatile loads: a fence on these paths
// (MEMBAR.CTA) would also wait for the warp's outstanding HBM loads and put an HBM round trip into the serial chain
// of every group.  Only the flusher, whose HBM stores must be visible before `flushed` moves, publishes with a
// release.  (mbarrier try_wait was measured for the hand-overs too: slower than polling here.)
__device__ __forceinline__ uint32_t ld_vol(uint32_t a) {
    uint32_t v;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void st_vol(uint32_t a, uint32_t v) { asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void st_release(uint32_t a, uint32_t v) { asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned long long ld_vol64(uint32_t a) {
    unsigned long long v;
    asm volatile("ld.volatile.shared.u64 %0, [%1];" : "=l"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void st_vol64(uint32_t a, unsigned long long v) { asm volatile("st.volatile.shared.u64 [%0], %1;" ::"r"(a), "l"(v) : "memory"); }
__device__ __forceinline__ uint32_t lds8(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts8(uint32_t a, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
// byte K of a batch, executed when rem > K (rem = bytes left for this lane, may be <= 0)
template <int K> __device__ __forceinline__ void lds8_if(uint32_t &v, uint32_t a, int rem) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.gt.s32 p, %2, %3;\n\t@p ld.shared.u8 %0, [%1+%3];\n\t}" : "=r"(v) : "r"(a), "r"(rem), "n"(K) : "memory");
}
template <int K> __device__ __forceinline__ void sts8_if(uint32_t a, uint32_t v, int rem) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.gt.s32 p, %2, %3;\n\t@p st.shared.u8 [%0+%3], %1;\n\t}" ::"r"(a), "r"(v), "r"(rem), "n"(K) : "memory");
}

// Waiting for the commit of group `ga - 1`: the warp next in line polls back to back, the others back off.
__device__ __forceinline__ void wait_turn(uint32_t done_g_addr, uint32_t ga) {
    for (;;) {
        const uint32_t d = ld_vol(done_g_addr);
        if (d == ga) return;
#ifdef LZB_XSLEEP
        if (ga - d > 2) __nanosleep(LZB_XSLEEP);
#endif
    }
}

struct XEnv {
    uint32_t ring_s;  // shared-window address of the ring
    uint32_t ctl_s;   // ... of the control block
    uint8_t *out_g;   // the stream's first output byte in HBM
    uint32_t bias;    // out_g & 15: ring offsets and HBM addresses agree modulo 16
    uint32_t lane;
    uint32_t total;   // bytes the stream's blocks announce: nobody ever writes a position at or beyond it
};

// One source byte at stream position p; `ring_from` = lowest position still intact in the ring.
__device__ __forceinline__ uint32_t x_read(const XEnv &e, uint32_t p, uint32_t ring_from) {
    if (p >= ring_from) return lds8(e.ring_s + ((p + e.bias) & kXMask));
    return __ldcg(e.out_g + p);
}

// Per-lane copies of up to kXSolo bytes, bounded by the warp-uniform `maxn`; the loads of a batch are issued before
// its stores (one shared-memory round trip per batch).  SRC_RING: source is the ring (a match), else the literal
// scratch in HBM.  doff / soff are ring offsets (bias applied, masked).
template <bool WRAP, bool SRC_RING>
__device__ __forceinline__ void solo_copy(uint32_t ring_s, uint32_t doff, uint32_t soff, const uint8_t *sg, uint32_t n, uint32_t maxn) {
#pragma unroll 1
    for (uint32_t g = 0; g < maxn; g += kXBatch) {
        const int rem = (int)n - (int)g;
        uint32_t t0, t1, t2, t3, t4, t5, t6, t7;
        if (!WRAP) {
            const uint32_t sa = ring_s + soff + g, da = ring_s + doff + g;
#define XLD(K, T) if (SRC_RING) lds8_if<K>(T, sa, rem); else if (rem > K) T = __ldg(sg + g + K);
            XLD(0, t0) XLD(1, t1) XLD(2, t2) XLD(3, t3) XLD(4, t4) XLD(5, t5) XLD(6, t6) XLD(7, t7)
#undef XLD
            sts8_if<0>(da, t0, rem); sts8_if<1>(da, t1, rem); sts8_if<2>(da, t2, rem); sts8_if<3>(da, t3, rem);
            sts8_if<4>(da, t4, rem); sts8_if<5>(da, t5, rem); sts8_if<6>(da, t6, rem); sts8_if<7>(da, t7, rem);
        } else {
#define XLD(K, T) if (SRC_RING) lds8_if<0>(T, ring_s + ((soff + g + K) & kXMask), rem - K); else if (rem > K) T = __ldg(sg + g + K);
            XLD(0, t0) XLD(1, t1) XLD(2, t2) XLD(3, t3) XLD(4, t4) XLD(5, t5) XLD(6, t6) XLD(7, t7)
#undef XLD
#define XST(K, T) sts8_if<0>(ring_s + ((doff + g + K) & kXMask), T, rem - K);
            XST(0, t0) XST(1, t1) XST(2, t2) XST(3, t3) XST(4, t4) XST(5, t5) XST(6, t6) XST(7, t7)
#undef XST
        }
    }
}

// The slow paths, kept out of line: short matches whose source reaches back beyond the ring (read from HBM, D > M)
// and long or short-period matches (the whole warp copies one at a time).  `mine` marks the lanes to serve; all
// lanes of the warp call.
__device__ __forceinline__ void x_special(XEnv e, bool mine, bool coop, uint32_t M, uint32_t D, uint32_t src, uint32_t my_dst, uint32_t ring_from) {
    const uint32_t lane = e.lane, bias = e.bias, ring_s = e.ring_s;
    if (__any_sync(kFull, mine && !coop)) {
        const uint32_t n = (mine && !coop) ? M : 0u;
        const uint32_t maxn = __reduce_max_sync(kFull, n);
#pragma unroll 1
        for (uint32_t g = 0; g < maxn; g += kXBatch) {
            uint32_t t[kXBatch];
#pragma unroll
            for (uint32_t k = 0; k < kXBatch; k++)
                if (g + k < n) t[k] = x_read(e, src + g + k, ring_from);
#pragma unroll
            for (uint32_t k = 0; k < kXBatch; k++)
                if (g + k < n) sts8(ring_s + ((my_dst + g + k + bias) & kXMask), t[k]);
        }
    }
    uint32_t cm = __ballot_sync(kFull, mine && coop);
    while (cm) {
        const int j = __ffs(cm) - 1;
        cm &= cm - 1;
        const uint32_t o = __shfl_sync(kFull, my_dst, j), d = __shfl_sync(kFull, D, j), n = __shfl_sync(kFull, M, j);
        if (d >= n) {
#pragma unroll 1
            for (uint32_t t0 = 0; t0 < n; t0 += 128) {  // four loads in flight per lane: the source may be in HBM
                uint32_t q[4];
#pragma unroll
                for (uint32_t k = 0; k < 4; k++)
                    if (t0 + k * 32 + lane < n) q[k] = x_read(e, o - d + t0 + k * 32 + lane, ring_from);
#pragma unroll
                for (uint32_t k = 0; k < 4; k++)
                    if (t0 + k * 32 + lane < n) sts8(ring_s + ((o + t0 + k * 32 + lane + bias) & kXMask), q[k]);
            }
        } else {  // byte i == byte (i mod D) of the D bytes before the match
#pragma unroll 1
            for (uint32_t t = lane; t < n; t += 32) sts8(ring_s + ((o + t + bias) & kXMask), x_read(e, o - d + t % d, ring_from));
        }
        __syncwarp();
    }
}
// Literal runs longer than kXSolo: the whole warp copies one at a time.
__device__ __forceinline__ void x_long_literals(XEnv e, uint32_t lm, uint32_t L, uint32_t my_out, const uint8_t *my_lit) {
    while (lm) {
        const int j = __ffs(lm) - 1;
        lm &= lm - 1;
        const uint32_t o = __shfl_sync(kFull, my_out, j), n = __shfl_sync(kFull, L, j);
        const uint8_t *s = reinterpret_cast<const uint8_t *>(__shfl_sync(kFull, reinterpret_cast<uintptr_t>(my_lit), j));
#pragma unroll 1
        for (uint32_t t = e.lane; t < n; t += 32) sts8(e.ring_s + ((o + t + e.bias) & kXMask), __ldg(s + t));
    }
}

// Literals and matches of the lanes marked `act`: a contiguous lane range whose output is [cur_base, range_end).
// Preconditions: the caller may write that range (ring discipline); if `have_turn`, every byte below cur_base is
// final, otherwise this function waits for the commit of group ga - 1 before it touches anything younger than the
// watermark it saw.  On return every byte of the range is final.
__device__ __forceinline__ void x_process(const XEnv &e, bool act, uint32_t L, uint32_t M, uint32_t D, uint32_t my_out, const uint8_t *my_lit,
                                          uint32_t cur_base, uint32_t range_end, bool have_turn, uint32_t ga) {
    const uint32_t lane = e.lane, bias = e.bias, ring_s = e.ring_s;
    const uint32_t my_dst = my_out + L;
    const bool dst_wrap = range_end > cur_base && (((cur_base + bias) ^ (range_end + bias - 1)) & ~kXMask) != 0;
    // ---- literals (fse_core.rs:112-118) ----
    {
        const uint32_t n = (act && L <= kXSolo) ? L : 0u;
        const uint32_t maxn = __reduce_max_sync(kFull, n);
        const uint32_t doff = (my_out + bias) & kXMask;
        if (!dst_wrap) solo_copy<false, false>(ring_s, doff, 0, my_lit, n, maxn);
        else solo_copy<true, false>(ring_s, doff, 0, my_lit, n, maxn);
        const uint32_t lm = __ballot_sync(kFull, act && L > kXSolo);
        if (lm) x_long_literals(e, lm, L, my_out, my_lit);
    }
    __syncwarp();
    // ---- matches (lz/writer.rs:144-180) ----
    // Round 0: matches whose source lies below the commit watermark seen now.  The others are assigned to rounds
    // 1.. by replaying "everything before the first unresolved match is final" on positions alone, so that the
    // part that has to wait for this group's turn is nothing but short copies.
    const uint32_t ring_lo = ld_vol(e.ctl_s + XOFF(ring_lo));
    // writers stay below min(cur_base + S, total), so what they can have overwritten lies W below that
    const uint32_t top = cur_base + kXAhead < e.total ? cur_base + kXAhead : e.total;
    const uint32_t keep = top > kXRing ? top - kXRing : 0u;
    const uint32_t ring_from = ring_lo > keep ? ring_lo : keep;
    const uint32_t src = my_dst - D;
    const uint32_t nonself_end = src + M < my_dst ? src + M : my_dst;
    const bool need = act && M != 0;
    const bool coop = M > kXSolo || (D < (uint32_t)kXBatch && D < M);  // short periods break the load batching
    const bool inring = src >= ring_from;
    const bool solo = need && !coop && inring;
    const uint32_t soff = (src + bias) & kXMask, doff = (my_dst + bias) & kXMask;
    const bool wrap = dst_wrap || __any_sync(kFull, solo && soff + M > kXRing);
    const uint32_t pd0 = have_turn ? cur_base : ld_vol(e.ctl_s + XOFF(pos_done));
    uint32_t my_round = (need && nonself_end > pd0) ? 0xFFu : 0u;
    uint32_t n_rounds = 1;
#pragma unroll 1
    for (uint32_t pend = __ballot_sync(kFull, my_round != 0); pend; n_rounds++) {
        const uint32_t wm = __shfl_sync(kFull, my_dst, __ffs(pend) - 1);
        const bool ready = ((pend >> lane) & 1u) && nonself_end <= wm;
        if (ready) my_round = n_rounds;
        pend &= ~__ballot_sync(kFull, ready);
    }
    const uint32_t max0 = __reduce_max_sync(kFull, (solo && my_round == 0) ? M : 0u);
    const uint32_t max1 = n_rounds > 1 ? __reduce_max_sync(kFull, (solo && my_round != 0) ? M : 0u) : 0u;
    const uint32_t special = __ballot_sync(kFull, need && !solo);  // lanes that take one of the slow paths
#pragma unroll 1
    for (uint32_t r = 0; r < n_rounds; r++) {
        if (r == 1 && !have_turn) wait_turn(e.ctl_s + XOFF(done_g), ga);
        const bool mine = need && my_round == r;
        {
            const uint32_t n = (mine && solo) ? M : 0u;
            const uint32_t maxn = r == 0 ? max0 : max1;
            if (!wrap) solo_copy<false, true>(ring_s, doff, soff, nullptr, n, maxn);
            else solo_copy<true, true>(ring_s, doff, soff, nullptr, n, maxn);
        }
        if (special && __any_sync(kFull, mine && !solo)) x_special(e, mine && !solo, coop, M, D, src, my_dst, ring_from);
        __syncwarp();
    }
    if (n_rounds == 1 && !have_turn) wait_turn(e.ctl_s + XOFF(done_g), ga);
}

// Copies stream bytes [a, b) from the ring to HBM; 16-byte units wherever the HBM address is aligned.
__device__ __forceinline__ void x_flush(const XEnv &e, uint32_t a, uint32_t b) {
    const uint32_t lane = e.lane, bias = e.bias;
    uint32_t head = (16u - ((a + bias) & 15u)) & 15u;
    if (head > b - a) head = b - a;
    if (lane < head) e.out_g[a + lane] = (uint8_t)lds8(e.ring_s + ((a + lane + bias) & kXMask));
    a += head;
    const uint32_t nv = (b - a) >> 4;
#pragma unroll 1
    for (uint32_t i = lane; i < nv; i += 32) {
        uint4 v;
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(e.ring_s + ((a + i * 16 + bias) & kXMask)) : "memory");
        __stcs(reinterpret_cast<uint4 *>(e.out_g + a + i * 16), v);
    }
    a += nv * 16;
    if (a + lane < b) e.out_g[a + lane] = (uint8_t)lds8(e.ring_s + ((a + lane + bias) & kXMask));
}

// Raw and LZVN blocks, produced directly in HBM (out of line: rare, and the LZVN interpreter is register hungry).
__device__ __noinline__ int x_plain_block(const uint8_t *blk, uint32_t type, uint32_t n_raw, uint64_t src_rest, uint8_t *out_g, uint32_t pos,
                                          uint64_t cap, uint32_t tid) {
    if (type == BT_RAW) {
        if (tid < 32) warp_copy(out_g + pos, blk + 8, n_raw, tid);
        return 0;
    }
    return tid == 0 ? vn_decode_block(blk, src_rest, out_g, pos, cap) : 0;
}

__global__ void __launch_bounds__(kXThreads, LZB_XCTAS)
k_expand_cta(const uint8_t *__restrict__ src_base, const uint64_t *__restrict__ src_off, const uint64_t *__restrict__ src_len,
             uint8_t *__restrict__ dst_base, const uint64_t *__restrict__ dst_off, const uint64_t *__restrict__ dst_cap,
             const StreamCounts *__restrict__ bases, const BlockDesc *__restrict__ blocks, const FseDesc *__restrict__ fse,
             const uint8_t *__restrict__ lit_scratch, const LmdRec *__restrict__ lmd_scratch, const uint64_t *__restrict__ raw_total,
             uint32_t *err, uint32_t n_streams, uint32_t *work_counter, const uint64_t *__restrict__ long_base) {
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool flusher = warp == kXWorkers;
    XEnv e;
    e.ring_s = (uint32_t)__cvta_generic_to_shared(x_smem);
    e.ctl_s = e.ring_s + kXRing;
    e.lane = lane;
    const uint32_t ctl_s = e.ctl_s;
    if (tid < kXSlots) st_vol64(ctl_s + XOFF(carry) + 8 * tid, 0ull);
    if (tid == 0) st_vol(ctl_s + XOFF(done_g), 0);
    uint32_t G0 = 0;  // groups this CTA has seen before the current block: numbering runs on across streams, every
                      // warp counts alike
    for (;;) {
        __syncthreads();  // the previous stream is completely done with the ring and the control block
        if (tid == 0) {
            st_vol(ctl_s + XOFF(stream), atomicAdd(work_counter, 1u));
            st_vol(ctl_s + XOFF(pos_done), 0); st_vol(ctl_s + XOFF(flushed), 0); st_vol(ctl_s + XOFF(ring_lo), 0); st_vol(ctl_s + XOFF(stop), 0);
        }
        __syncthreads();
        const uint32_t stream = ld_vol(ctl_s + XOFF(stream));
        if (stream >= n_streams) return;
        e.out_g = dst_base + dst_off[stream];
        e.bias = (uint32_t)(reinterpret_cast<uintptr_t>(e.out_g) & 15u);
        e.total = (uint32_t)raw_total[stream];
        const uint64_t b0 = bases[stream].n_blocks;
        uint64_t b1 = bases[stream + 1].n_blocks;
        if (b1 - b0 == 1 && vn_fast_eligible(1, blocks[b0], src_off[stream] + src_len[stream] - blocks[b0].src_off, dst_cap[stream])) b1 = b0;  // k_expand_vn
        if (long_base[stream] != ~0ull) b1 = b0;  // expand_long.cu
        uint32_t fl = 0;   // flusher's copy of ctl->flushed
        for (uint64_t b = b0; b < b1; b++) {
            const BlockDesc bd = blocks[b];
            const uint32_t pos = (uint32_t)(bd.dst_off - dst_off[stream]);
            if (bd.type == BT_VX1 || bd.type == BT_VX2) {
                const FseDesc *fd = fse + bd.fse_idx;
                if (!(fd->ok_lit && fd->ok_lmd)) break;  // the entropy stages already recorded why
                const uint32_t n_lmds = fd->n_lmds;
                const uint32_t block_end = pos + fd->n_raw;
                const uint32_t ng = (n_lmds + 31) >> 5;
                if (flusher) {
                    for (;;) {
                        const uint32_t pd = ld_vol(ctl_s + XOFF(pos_done));
                        const uint32_t top = (pd + e.bias) & ~15u;
                        const uint32_t lim = top > e.bias ? top - e.bias : 0u;
                        if (lim > fl && (lim - fl >= kXFlushMin || pd >= block_end)) {
                            x_flush(e, fl, lim);
                            fl = lim;
                            __syncwarp();
                            if (lane == 0) st_release(ctl_s + XOFF(flushed), fl);
                        } else if (pd < block_end) {
                            __nanosleep(200);
                        }
                        if (pd >= block_end) break;
                    }
                    G0 += ng;
                    continue;
                }
                const uint8_t *lit = lit_scratch + fd->lit_off;
                const uint2 *recs = reinterpret_cast<const uint2 *>(lmd_scratch + fd->lmd_off);
                uint32_t lg = (warp + kXWorkers - G0 % kXWorkers) % kXWorkers;
                uint2 nxt = make_uint2(0, 0);
                if (lg * 32 + lane < n_lmds) nxt = __ldg(recs + lg * 32 + lane);
#pragma unroll 1
                for (; lg < ng; lg += kXWorkers) {
                    const uint2 rec = nxt;
                    nxt = make_uint2(0, 0);
                    if ((lg + kXWorkers) * 32 + lane < n_lmds) nxt = __ldg(recs + (lg + kXWorkers) * 32 + lane);
                    const uint32_t ga = G0 + lg;
                    const uint32_t L = rec.x & 0xFFFF, M = rec.x >> 16, D = rec.y;
                    // inclusive scan of (sum L) << 17 | (sum L+M): 32*315 < 2^14, 32*(315+2359) < 2^17
                    const uint32_t v = (L << 17) + (L + M);
                    uint32_t inc = v;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const uint32_t t = __shfl_up_sync(kFull, inc, o);
                        if (lane >= (uint32_t)o) inc += t;
                    }
                    const uint32_t exc = inc - v, tot = __shfl_sync(kFull, inc, 31);
                    const uint32_t tot_o = tot & 0x1FFFF, tot_l = tot >> 17;
                    // chained scan: my base = my predecessor's end
                    uint32_t gb = pos, lb = 0;
                    if (lg != 0) {
                        const uint32_t slot = ctl_s + XOFF(carry) + 8 * ((ga - 1) & (kXSlots - 1));
                        const uint32_t want = (ga & 0x7FFF) | 0x8000;
                        unsigned long long c;
                        do { c = ld_vol64(slot); } while (((uint32_t)c & 0xFFFF) != want);
                        gb = (uint32_t)(c >> 32);
                        lb = ((uint32_t)c >> 16) & 0xFFFF;
                    }
                    const uint32_t ge = gb + tot_o;
                    if (lane == 0) {
                        st_vol64(ctl_s + XOFF(carry) + 8 * (ga & (kXSlots - 1)),
                                 ((unsigned long long)ge << 32) | ((unsigned long long)(lb + tot_l) << 16) | (((ga + 1) & 0x7FFF) | 0x8000));
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(lit + lb + tot_l + 384));  // scratch has slack past its end
                    }
                    const uint32_t my_out = gb + (exc & 0x1FFFF);
                    const uint8_t *my_lit = lit + lb + (exc >> 17);
                    // A group with a lot of output cannot run ahead of the flusher: it takes its turn first and is
                    // then committed in sub-ranges of at most kXBig bytes (one LMD is at most 2674 bytes).
                    const bool big = tot_o > kXBig;
                    if (big) wait_turn(ctl_s + XOFF(done_g), ga);
                    uint32_t a = 0;
                    do {
                        uint32_t start = gb, sub_end = ge, bnd = 32;
                        if (big) {
                            start = __shfl_sync(kFull, my_out, a);
                            const bool fits = lane >= a && my_out + L + M - start <= kXBig;
                            const uint32_t nf = ~(__ballot_sync(kFull, fits) >> a);
                            const uint32_t cnt = nf ? (uint32_t)__ffs(nf) - 1u : 32u;
                            bnd = a + cnt < 32 ? a + cnt : 32;
                            sub_end = __shfl_sync(kFull, my_out + L + M, bnd - 1);
                        }
                        while (sub_end > ld_vol(ctl_s + XOFF(flushed)) + kXAhead) __nanosleep(32);
                        x_process(e, lane >= a && lane < bnd, L, M, D, my_out, my_lit, start, sub_end, big, ga);
                        if (lane == 0) st_vol(ctl_s + XOFF(pos_done), sub_end);
                        __syncwarp();
                        a = bnd;
                    } while (a < 32);
                    if (lane == 0) st_vol(ctl_s + XOFF(done_g), ga + 1);
                }
                G0 += ng;
            } else {
                // Raw and LZVN blocks bypass the ring: drain it, produce the block in HBM, restart the ring after it.
                __syncthreads();
                if (flusher) x_flush(e, fl, ld_vol(ctl_s + XOFF(pos_done)));
                __syncthreads();
                const int st = x_plain_block(src_base + bd.src_off, bd.type, bd.n_raw, src_off[stream] + src_len[stream] - bd.src_off, e.out_g, pos,
                                             dst_cap[stream], tid);
                if (st) {
                    const uint32_t kb = bd.index < 0x1FFFFFu ? bd.index : 0x1FFFFFu;
                    atomicMin(&err[stream], err_key(kb, PH_LMD, st));
                    st_vol(ctl_s + XOFF(stop), 1);
                }
                __syncthreads();
                if (ld_vol(ctl_s + XOFF(stop))) break;
                fl = pos + bd.n_raw;
                if (tid == 0) { st_vol(ctl_s + XOFF(pos_done), fl); st_vol(ctl_s + XOFF(flushed), fl); st_vol(ctl_s + XOFF(ring_lo), fl); }
                __syncthreads();
            }
        }
        __syncthreads();  // every group is committed
        if (flusher) x_flush(e, fl, ld_vol(ctl_s + XOFF(pos_done)));
    }
}

int setup_expand_kernel() {
    return (int)cudaFuncSetAttribute(k_expand_cta, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kXSmem);
}

void launch_expand_cta(const uint8_t *src, const uint64_t *src_off, const uint64_t *src_len, uint8_t *dst, const uint64_t *dst_off,
                       const uint64_t *dst_cap, const StreamCounts *bases, const BlockDesc *blocks, const FseDesc *fse, const uint8_t *lit_scratch,
                       const LmdRec *lmd_scratch, const uint64_t *raw_total, uint32_t *err, size_t n, uint32_t *work_counter /* zeroed */,
                       const uint64_t *long_base, int n_sms, cudaStream_t s) {
    if (n == 0) return;
    const size_t resident = (size_t)n_sms * LZB_XCTAS;
    const unsigned grid = (unsigned)(n < resident ? n : resident);
    k_expand_cta<<<grid, kXThreads, kXSmem, s>>>(src, src_off, src_len, dst, dst_off, dst_cap, bases, blocks, fse, lit_scratch, lmd_scratch, raw_total,
                                                 err, (uint32_t)n, work_counter, long_base);
}

}  // namespace lzb
