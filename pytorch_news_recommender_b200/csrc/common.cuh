// common.cuh — shared device helpers for libnrms_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#ifndef __CUDA_ARCH__
#define NRMS_HOST 1
#endif

namespace nrms {

constexpr int kWarp = 32;
constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

__host__ __device__ inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline int64_t align_up(int64_t a, int64_t b) { return ceil_div64(a, b) * b; }

// ---------------------------------------------------------------------------------------
// Philox4x32-10 counter RNG (128 random bits per call) keyed by the step's dropout seed.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0;
        key.y += W1;
    }
    return ctr;
}

// Dropout masks are addressed by (stream, row, 8-column group): one Philox call yields the 8
// keep decisions of columns [8g, 8g+8) of a row (16 random bits per element), so a kernel that
// owns 8 consecutive columns pays one call, and the keep bits are stored (1 bit per element,
// byte g of a row = group g) for the backward instead of being regenerated.
struct Dropout {
    uint32_t thresh;  // drop when the 16 random bits < thresh ; 0 => dropout disabled
    float scale;      // 1/(1-p)
    uint32_t seed_lo, seed_hi;

    __host__ __device__ bool enabled() const { return thresh != 0u; }

    // 8 keep bits (bit j = column 8g+j is kept) of group g of row `row` in stream `sid`
    __device__ __forceinline__ uint32_t keep8(uint32_t sid, uint64_t row, uint32_t g) const {
        const uint4 r = philox4x32_10(make_uint4((uint32_t)row, (uint32_t)(row >> 32), g, sid),
                                      make_uint2(seed_lo, seed_hi));
        const uint32_t w[4] = {r.x, r.y, r.z, r.w};
        uint32_t bits = 0u;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint32_t v = (w[j >> 1] >> (16 * (j & 1))) & 0xffffu;
            bits |= (v >= thresh ? 1u : 0u) << j;
        }
        return bits;
    }
};

inline Dropout make_dropout(float p, uint64_t seed) {
    Dropout d;
    if (p <= 0.f) {
        d.thresh = 0u;
        d.scale = 1.f;
    } else {
        double t = (double)p * 65536.0;
        if (t > 65535.0) t = 65535.0;
        if (t < 1.0) t = 1.0;
        d.thresh = (uint32_t)t;
        d.scale = 1.f / (1.f - p);
    }
    d.seed_lo = (uint32_t)seed;
    d.seed_hi = (uint32_t)(seed >> 32);
    return d;
}

constexpr uint32_t kDropEmbedding = 1u;  // nrms_v0.py:137
constexpr uint32_t kDropContext = 2u;    // nrms_v0.py:171-173

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

}  // namespace nrms
