// encode.cu -- batched LZFSE encode kernels for sm_100a (placeholder: entry points only).
#include <new>
#include <string>

#include "common.cuh"
#include "host_util.h"

using namespace lzb;

struct lzfse_b200_encoder {
    int device = 0;
    std::string last_error;
    uint64_t launches = 0;
};

extern "C" {
int lzfse_b200_encoder_create(int, lzfse_b200_encoder **out) { if (out) *out = nullptr; return LZFSE_B200_NO_DEVICE; }
void lzfse_b200_encoder_destroy(lzfse_b200_encoder *e) { delete e; }
const char *lzfse_b200_encoder_last_error(const lzfse_b200_encoder *e) { return e ? e->last_error.c_str() : ""; }
uint64_t lzfse_b200_encoder_last_launches(const lzfse_b200_encoder *e) { return e ? e->launches : 0; }
size_t lzfse_b200_encode_bound(size_t n) { return n + n / 4 + (n / 16384 + 2) * 768 + 64; }
int lzfse_b200_encode_bytes(lzfse_b200_encoder *, const uint8_t *, size_t, uint8_t *, size_t, size_t *) { return LZFSE_B200_INVALID_ARGUMENT; }
int lzfse_b200_encode_batch_device(lzfse_b200_encoder *, const uint8_t *, const uint64_t *, const uint64_t *, uint8_t *, const uint64_t *,
                                   const uint64_t *, uint64_t *, int32_t *, size_t, void *) { return LZFSE_B200_INVALID_ARGUMENT; }
int lzfse_b200_encode_batch_host(lzfse_b200_encoder *, const uint8_t *, const uint64_t *, const uint64_t *, uint8_t *, const uint64_t *,
                                 const uint64_t *, uint64_t *, int32_t *, size_t) { return LZFSE_B200_INVALID_ARGUMENT; }
}
