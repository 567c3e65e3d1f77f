// common.cuh -- shared definitions for the sm_100a LZFSE kernels.
//
// Wire-format constants restate lzfse_rust v0.2.0 (paths relative to its src/):
//   fse/constants.rs:22-69 (block limits, symbol/state counts, header sizes)
//   base/magic_bytes.rs:3-7 (block magics), vn/constants.rs:1-13, encode/constants.rs:3-10.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/lzfse_b200.h"

namespace lzb {

constexpr uint32_t kLmdsPerBlock = 10000;
constexpr uint32_t kLiteralsPerBlock = 40000;
constexpr uint32_t kLStates = 64, kMStates = 64, kDStates = 256, kUStates = 1024;
constexpr uint32_t kLSymbols = 20, kMSymbols = 20, kDSymbols = 64, kUSymbols = 256;
constexpr uint32_t kNWeights = 360;
constexpr uint32_t kMaxLValue = 315, kMaxMValue = 2359, kMaxDValue = 262139;
constexpr uint32_t kV1HeaderSize = 50, kV2HeaderSize = 32;
constexpr uint32_t kV1WeightBytes = 722, kV2WeightBytesMax = 630;
constexpr uint32_t kMaxLBits = 14, kMaxMBits = 17, kMaxDBits = 23, kMaxUBits = 10;

constexpr uint32_t kMagicEos = 0x24787662u, kMagicRaw = 0x2D787662u, kMagicVx1 = 0x31787662u,
                   kMagicVx2 = 0x32787662u, kMagicVxn = 0x6E787662u;
constexpr uint32_t kVnHeaderSize = 12, kVnPayloadLimit = 0x2000, kVnMaxD = 65535;

constexpr uint64_t kMaxStreamRaw = 0xF0000000ull;  // decoded bytes per stream (32-bit positions in expand.cu)

constexpr uint32_t kGoodMatchLen = 40, kRawCutoff = 20, kRawLimit = 0x4000, kVnCutoff = 4096;
constexpr uint32_t kHashBits = 14, kHashWidth = 4;

enum BlockType : uint32_t { BT_RAW = 0, BT_VXN = 1, BT_VX1 = 2, BT_VX2 = 3 };

// Error keys order the failures of one stream the way the reference's sequential decoder meets
// them (decode/decoder.rs:72-141): block index first, then the stage inside the block.
enum Phase : uint32_t { PH_HEADER = 0, PH_WEIGHTS = 1, PH_LIT_TAKE = 2, PH_LIT = 3, PH_LMD_TAKE = 4, PH_LMD = 5 };
constexpr uint32_t kNoError = 0xFFFFFFFFu;
// Work counters of one decode chain (zeroed per call): 0 literals (exact kernel), 1 LMDs, 2 expansion, 3 LZVN expansion,
// 4-7 spare;
// long streams (expand_long.cu): 8-9 image elements (u64), 10 blocks, 11 streams as counted by k_scan<false>;
// 12-13 / 14 / 15 the same three as allocation cursors of k_scan<true>; 16 pass-1 work counter
constexpr uint32_t kWorkWords = 32;
struct LongTotals { uint64_t elements; uint32_t n_blocks, n_streams; };  // follows the StreamCounts totals in pinned host memory
__host__ __device__ inline uint32_t err_key(uint32_t block, uint32_t phase, uint32_t code) {
    return (block << 11) | (phase << 8) | code;
}

// One compressed block of one stream (filled by the frame scan).
struct BlockDesc {
    uint64_t src_off;  // absolute byte offset of the block's magic inside src_base
    uint64_t dst_off;  // absolute byte offset of the block's first output byte inside dst_base
    uint32_t n_raw;    // header n_raw_bytes
    uint32_t stream;
    uint32_t type;     // BlockType
    uint32_t index;    // block index inside its stream
    uint32_t fse_idx;  // index into FseDesc[] for bvx1/bvx2
    uint32_t pad;      // 1: LZVN block whose n_raw was clamped to the destination by the scan (see k_scan)
};

// Parsed bvx1/bvx2 header plus scratch placement (fse/block.rs:80-136).
struct FseDesc {
    uint64_t lit_off;  // byte offset into the literal scratch
    uint64_t lmd_off;  // element offset into the LMD scratch
    uint32_t block;    // BlockDesc index
    uint32_t flags;    // bit0 v1 header, bit1 literal payload truncated, bit2 lmd payload truncated
    uint32_t header_size;  // bytes from block start to the literal payload (header + weights)
    uint32_t n_weight_bytes;
    uint32_t n_literals, n_lit_payload, lit_bits;
    uint32_t n_lmds, n_lmd_payload, lmd_bits;
    uint16_t lit_state[4];
    uint16_t lmd_state[3];
    uint16_t pad;
    uint32_t n_raw;
    uint32_t ok_lit;   // 1 once the literal stage validated and produced this block's literals
    uint32_t ok_lmd;   // 1 once the LMD stage validated and produced this block's LMD list
    uint32_t pad2;
};
constexpr uint32_t FSE_V1 = 1, FSE_TRUNC_LIT = 2, FSE_TRUNC_LMD = 4;

// Per-stream counters; exclusive-scanned in place into per-stream bases.
struct StreamCounts {
    uint64_t n_blocks, n_fse, n_literals, n_lmds;
};

// One decoded LMD as the expansion stage consumes it: D already substituted (lmd/lmd_type.rs:155-159).
struct __align__(8) LmdRec {
    uint16_t l, m;
    uint32_t d;
};

__device__ __forceinline__ uint32_t ld_u8(const uint8_t *p) { return *p; }
__device__ __forceinline__ uint32_t ld_u16(const uint8_t *p) { return ld_u8(p) | (ld_u8(p + 1) << 8); }
__device__ __forceinline__ uint32_t ld_u32(const uint8_t *p) {
    if ((reinterpret_cast<uintptr_t>(p) & 3) == 0) return *reinterpret_cast<const uint32_t *>(p);
    return ld_u8(p) | (ld_u8(p + 1) << 8) | (ld_u8(p + 2) << 16) | (ld_u8(p + 3) << 24);
}
__device__ __forceinline__ uint64_t ld_u64(const uint8_t *p) { return (uint64_t)ld_u32(p) | ((uint64_t)ld_u32(p + 4) << 32); }
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

}  // namespace lzb
