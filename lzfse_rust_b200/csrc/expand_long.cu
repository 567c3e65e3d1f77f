// expand_long.cu -- LZ expansion of LONG streams (many bvx blocks, matches reaching up to 262 139 bytes back across
// block boundaries): two passes, so that the blocks of ONE stream expand side by side.
//
// The reference walks a stream's blocks strictly in order (decode/decoder.rs:72-99): block b's matches read what
// blocks b-1, b-2, ... wrote (lz/writer.rs:144-180).  With one warp or one CTA per stream that order is the whole run
// time -- 8 streams of 16 MiB kept 8 SMs busy and decoded slower than 8 host threads.  An in-order formulation cannot
// spread a stream's blocks over the machine either: a match that reaches D bytes back with D below the block size lands
// in the part of the previous block that a sibling working in lock step has not produced yet, so the siblings fall into
// single file.  What does parallelise is keeping the VALUE of a byte apart from WHERE IT COMES FROM:
//
//   pass 1 (k_expand_long_p1, a warp per block, all blocks of all long streams at once) expands every block on its own
//          into a 32-bit image of its output: an element is either a final byte (0xFF000000 | b: a literal, or a copy of
//          one inside the block) or the stream position of the byte it equals, whenever the LMD chain leaves the block.
//          Copies inside the block copy elements, so a chain of matches ends in a literal or in ONE outside position.
//   pass 2 (k_expand_long_p2, a thread-block CLUSTER per stream, blocks in order) turns the image into bytes: final
//          elements are stored, the others are one gather from the stream's earlier output -- which is final, because
//          the blocks before were finished first.  A block's gather is fully parallel; only the block order is serial,
//          one cluster barrier per block.  (One CTA per stream was bound by what a single SM can gather: a scattered
//          byte load occupies its L1 for a cycle, 64 Ki of them per block, 27 us per block measured; the eight SMs of a
//          cluster share that.)
//
// Raw and LZVN blocks inside a long stream are produced by pass 2 when their turn comes (the reference's encoder never
// emits them there; Apple's may).  Errors: the entropy stages validated every distance against the bytes the stream has
// produced before the match (fse_core.rs:104-131), so pass 1 only touches blocks both stages accepted, and pass 2 stops at
// the first block they did not -- the reference's behaviour.
#include <cooperative_groups.h>
#include <cstdlib>

#include "common.cuh"
#include "lz_blocks.cuh"

namespace cg = cooperative_groups;

namespace lzb {

constexpr uint32_t kFinal = 0xFF000000u;  // element tags: positions are < kMaxStreamRaw = 0xF0000000
constexpr uint32_t kFull = 0xFFFFFFFFu;
constexpr uint32_t kP1Solo = 32;          // per-lane copies up to this many elements; longer ones go warp-wide
constexpr int kP1Warps = 8;
constexpr int kP2Threads = 1024;

// One bvx1/bvx2 block into its image V (V[0] = the block's first output byte); `pos` = that byte's stream position.
__device__ void p1_block(uint32_t *__restrict__ V, uint32_t pos, const uint8_t *__restrict__ lit, const LmdRec *__restrict__ lmds, uint32_t n_lmds,
                         uint32_t lane) {
    uint32_t out_base = 0, lit_base = 0;
    uint2 nxt = make_uint2(0, 0);
    if (lane < n_lmds) nxt = __ldg(reinterpret_cast<const uint2 *>(lmds) + lane);
    for (uint32_t b = 0; b < n_lmds; b += 32) {
        const uint2 rec = nxt;
        nxt = make_uint2(0, 0);
        if (b + 32 + lane < n_lmds) nxt = __ldg(reinterpret_cast<const uint2 *>(lmds) + b + 32 + lane);
        const uint32_t L = rec.x & 0xFFFF, M = rec.x >> 16, D = rec.y;
        // inclusive scan of (sum L) << 17 | (sum L+M): 32*315 < 2^14, 32*(315+2359) < 2^17
        uint32_t v = (L << 17) + (L + M), inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(kFull, inc, o);
            if (lane >= (uint32_t)o) inc += t;
        }
        const uint32_t exc = inc - v, tot = __shfl_sync(kFull, inc, 31);
        const uint32_t my_out = out_base + (exc & 0x1FFFF), my_lit = lit_base + (exc >> 17);
        // ---- literals: final elements ----
        {
            const uint32_t max_l = __reduce_max_sync(kFull, L);
            const uint8_t *ps = lit + my_lit;
            uint32_t *pd = V + my_out;
            for (uint32_t g = 0; g < max_l; g += 4) {
                uint32_t t[4];
#pragma unroll
                for (uint32_t k = 0; k < 4; k++)
                    if (g + k < L) t[k] = __ldg(ps + g + k);
#pragma unroll
                for (uint32_t k = 0; k < 4; k++)
                    if (g + k < L) pd[g + k] = kFinal | t[k];
            }
        }
        // ---- matches ----
        // Source element u of a match: inside the block (u >= 0) it is whatever the image holds there, outside it is
        // the position itself.  A lane copies on its own when everything it reads was settled before this step.
        const uint32_t my_dst = my_out + L;
        const int32_t src = (int32_t)my_dst - (int32_t)D;  // block-relative, may be negative
        const int32_t nonself_end = src + (int32_t)M < (int32_t)my_dst ? src + (int32_t)M : (int32_t)my_dst;
        const bool solo = M != 0 && M <= kP1Solo && D >= M && nonself_end <= (int32_t)out_base;
        {
            const uint32_t sm = solo ? M : 0u;
            const uint32_t max_m = __reduce_max_sync(kFull, sm);
            for (uint32_t g = 0; g < max_m; g += 4) {
                uint32_t t[4];
#pragma unroll
                for (uint32_t k = 0; k < 4; k++)
                    if (g + k < sm) {
                        const int32_t u = src + (int32_t)(g + k);
                        t[k] = u < 0 ? pos + (uint32_t)u : V[u];
                    }
#pragma unroll
                for (uint32_t k = 0; k < 4; k++)
                    if (g + k < sm) V[my_dst + g + k] = t[k];
            }
        }
        __syncwarp();
        // the rest in order, the whole warp on one match: element t equals element (t mod D) of the D elements before it
        uint32_t mask = __ballot_sync(kFull, M != 0 && !solo);
        while (mask) {
            const int j = __ffs(mask) - 1;
            mask &= mask - 1;
            const uint32_t o = __shfl_sync(kFull, my_dst, j), d = __shfl_sync(kFull, D, j), n = __shfl_sync(kFull, M, j);
            const int32_t s0 = (int32_t)o - (int32_t)d;
            for (uint32_t t = lane; t < n; t += 32) {
                const int32_t u = s0 + (int32_t)(d < n ? t % d : t);
                V[o + t] = u < 0 ? pos + (uint32_t)u : V[u];
            }
            __syncwarp();
        }
        out_base += tot & 0x1FFFF;
        lit_base += tot >> 17;
    }
}

__global__ void __launch_bounds__(kP1Warps * 32, 4)
k_expand_long_p1(const uint64_t *__restrict__ dst_off, const BlockDesc *__restrict__ blocks, const FseDesc *__restrict__ fse,
                 const uint8_t *__restrict__ lit_scratch, const LmdRec *__restrict__ lmd_scratch, const uint32_t *__restrict__ long_blocks,
                 const uint32_t *__restrict__ n_long_blocks, const uint64_t *__restrict__ long_base, uint32_t *__restrict__ image,
                 uint32_t *work_counter) {
    const uint32_t lane = lane_id();
    const uint32_t n = *n_long_blocks;
    for (;;) {
        uint32_t item = 0;
        if (lane == 0) item = atomicAdd(work_counter, 1u);
        item = __shfl_sync(kFull, item, 0);
        if (item >= n) return;
        const BlockDesc bd = blocks[long_blocks[item]];
        if (bd.type != BT_VX1 && bd.type != BT_VX2) continue;
        const FseDesc &fd = fse[bd.fse_idx];
        if (!(fd.ok_lit && fd.ok_lmd)) continue;  // pass 2 stops there
        const uint32_t pos = (uint32_t)(bd.dst_off - dst_off[bd.stream]);
        p1_block(image + long_base[bd.stream] + pos, pos, lit_scratch + fd.lit_off, lmd_scratch + fd.lmd_off, fd.n_lmds, lane);
        __syncwarp();
    }
}

// n image elements -> bytes at O; S = the stream's first output byte.  Everything a non-final element points at was
// written by this CTA before the barrier that precedes the call; the gather reads it from L2 (the L1 may hold the
// line from before).
__device__ __forceinline__ uint32_t p2_byte(uint32_t x, const uint8_t *S) { return x >= kFinal ? (x & 0xFFu) : (uint32_t)__ldcg(S + x); }

__device__ void p2_gather(const uint32_t *__restrict__ V, uint8_t *__restrict__ O, const uint8_t *S, uint32_t n, uint32_t tid) {
    uint32_t head = (4u - ((uint32_t)(reinterpret_cast<uintptr_t>(V) >> 2) & 3u)) & 3u;  // elements up to the first 16-byte boundary of the image
    if (head > n) head = n;
    if (tid < head) O[tid] = (uint8_t)p2_byte(__ldcs(V + tid), S);
    V += head; O += head; n -= head;
    const uint32_t n4 = n >> 2;
    const uint4 *V4 = reinterpret_cast<const uint4 *>(V);
    const bool aligned = (reinterpret_cast<uintptr_t>(O) & 3u) == 0;
    for (uint32_t j = tid; j < n4; j += 4 * kP2Threads) {  // four independent 16-byte loads, then their sixteen gathers
        uint4 x[4];
#pragma unroll
        for (uint32_t k = 0; k < 4; k++)
            if (j + k * kP2Threads < n4) x[k] = __ldcs(V4 + j + k * kP2Threads);
        uint32_t w[4];
#pragma unroll
        for (uint32_t k = 0; k < 4; k++)
            if (j + k * kP2Threads < n4)
                w[k] = p2_byte(x[k].x, S) | (p2_byte(x[k].y, S) << 8) | (p2_byte(x[k].z, S) << 16) | (p2_byte(x[k].w, S) << 24);
#pragma unroll
        for (uint32_t k = 0; k < 4; k++)
            if (j + k * kP2Threads < n4) {
                uint8_t *o = O + 4 * (size_t)(j + k * kP2Threads);
                if (aligned) *reinterpret_cast<uint32_t *>(o) = w[k];
                else { o[0] = (uint8_t)w[k]; o[1] = (uint8_t)(w[k] >> 8); o[2] = (uint8_t)(w[k] >> 16); o[3] = (uint8_t)(w[k] >> 24); }
            }
    }
    const uint32_t done = n4 * 4;
    if (tid < n - done) O[done + tid] = (uint8_t)p2_byte(__ldcs(V + done + tid), S);
}

// CTAs (SMs) per stream: the cluster size chosen at launch (8, see p2_cluster_size).
__global__ void __launch_bounds__(kP2Threads, 1)
k_expand_long_p2(const uint8_t *__restrict__ src_base, const uint64_t *__restrict__ src_off, const uint64_t *__restrict__ src_len,
                 uint8_t *__restrict__ dst_base, const uint64_t *__restrict__ dst_off, const uint64_t *__restrict__ dst_cap,
                 const StreamCounts *__restrict__ bases, const BlockDesc *__restrict__ blocks, const FseDesc *__restrict__ fse,
                 const uint32_t *__restrict__ long_streams, const uint64_t *__restrict__ long_base, const uint32_t *__restrict__ image,
                 uint32_t *err) {
    __shared__ int stop;  // rank 0's copy is the cluster's (read through distributed shared memory)
    cg::cluster_group cluster = cg::this_cluster();
    const uint32_t tid = threadIdx.x, rank = cluster.block_rank(), kP2Cluster = cluster.num_blocks();
    const uint32_t stream = long_streams[blockIdx.x / kP2Cluster];
    uint8_t *S = dst_base + dst_off[stream];
    const uint32_t *img = image + long_base[stream];
    if (tid == 0) stop = 0;
    cluster.sync();
    const int *stop0 = cluster.map_shared_rank(&stop, 0);
    const uint64_t b0 = bases[stream].n_blocks, b1 = bases[stream + 1].n_blocks;
    for (uint64_t b = b0; b < b1; b++) {
        const BlockDesc bd = blocks[b];
        const uint32_t pos = (uint32_t)(bd.dst_off - dst_off[stream]);
        if (bd.type == BT_VX1 || bd.type == BT_VX2) {
            const FseDesc &fd = fse[bd.fse_idx];
            if (!(fd.ok_lit && fd.ok_lmd)) break;  // the entropy stages already recorded why (same for every thread of the cluster)
            // this CTA's slice of the block: a multiple of 16 elements, so the slices keep the image's alignment
            const uint32_t per = ((fd.n_raw + kP2Cluster - 1) / kP2Cluster + 15u) & ~15u;
            const uint32_t lo = rank * per < fd.n_raw ? rank * per : fd.n_raw, hi = lo + per < fd.n_raw ? lo + per : fd.n_raw;
            p2_gather(img + pos + lo, S + pos + lo, S, hi - lo, tid);
        } else if (rank == 0) {
            if (bd.type == BT_RAW) {
                if (tid < 32) warp_copy(S + pos, src_base + bd.src_off + 8, bd.n_raw, tid);
            } else if (tid == 0) {
                const int st = vn_decode_block(src_base + bd.src_off, src_off[stream] + src_len[stream] - bd.src_off, S, pos, dst_cap[stream]);
                if (st) {
                    const uint32_t kb = bd.index < 0x1FFFFFu ? bd.index : 0x1FFFFFu;
                    atomicMin(&err[stream], err_key(kb, PH_LMD, st));
                    stop = 1;
                }
            }
        }
        cluster.sync();  // the block's bytes are final (and visible) for everything after it
        if (*stop0) break;
    }
    cluster.sync();  // nobody leaves while a sibling may still read its shared memory
}

// Cluster size of pass 2.  8 (the portable maximum) is the default: with 16-CTA clusters the 8 x 16 MiB configuration expands in
// 3.28 ms instead of 2.36 -- the barrier over 16 SMs costs more than halving the gather saves.  LZB_P2_CLUSTER overrides it for
// measurements.
static int p2_cluster_size() {
    static const int size = [] {
        const bool big_ok = cudaFuncSetAttribute(k_expand_long_p2, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess;
        if (!big_ok) cudaGetLastError();
        if (const char *e = getenv("LZB_P2_CLUSTER")) {
            const int v = atoi(e);
            if (v == 1 || v == 2 || v == 4 || v == 8 || (v == 16 && big_ok)) return v;
        }
        return 8;
    }();
    return size;
}

int launch_expand_long(const uint8_t *src, const uint64_t *src_off, const uint64_t *src_len, uint8_t *dst, const uint64_t *dst_off,
                       const uint64_t *dst_cap, const StreamCounts *bases, const BlockDesc *blocks, const FseDesc *fse, const uint8_t *lit_scratch,
                       const LmdRec *lmd_scratch, const uint32_t *long_blocks, const uint32_t *long_streams, const uint64_t *long_base,
                       uint32_t *image, uint32_t *err, uint32_t n_long_streams, uint32_t n_long_blocks, uint32_t *work /* kWorkWords */, int n_sms,
                       cudaStream_t s) {
    if (n_long_streams == 0) return 0;
    const unsigned need = (n_long_blocks + kP1Warps - 1) / kP1Warps, resident = (unsigned)n_sms * 4;
    k_expand_long_p1<<<need < resident ? need : resident, kP1Warps * 32, 0, s>>>(dst_off, blocks, fse, lit_scratch, lmd_scratch, long_blocks, work + 14, long_base,
                                                                                 image, work + 16);
    const int cl = p2_cluster_size();
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(n_long_streams * (unsigned)cl); cfg.blockDim = dim3(kP2Threads); cfg.dynamicSmemBytes = 0; cfg.stream = s;
    cudaLaunchAttribute at;
    at.id = cudaLaunchAttributeClusterDimension; at.val.clusterDim.x = (unsigned)cl; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
    cfg.attrs = &at; cfg.numAttrs = 1;
    return (int)cudaLaunchKernelEx(&cfg, k_expand_long_p2, src, src_off, src_len, dst, dst_off, dst_cap, bases, blocks, fse, long_streams, long_base,
                                   (const uint32_t *)image, err);
}

}  // namespace lzb
