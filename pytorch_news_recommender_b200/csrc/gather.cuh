// gather.cuh — word-embedding gather + embedding dropout (NewsEncoder.forward nrms_v0.py:166:
// `F.dropout(self.word_embedding(news))`), producing the operand formats of the projections.
//
// HBM-bound: a thread owns 8 consecutive columns of a token row (two float4 reads of the table
// row, one Philox call for its 8 dropout decisions; consecutive lanes take consecutive groups, so
// reads and writes of a row are coalesced) and writes
//   * the split-bf16 image unit (16 B to each plane; a row of a chunk is one 128-byte line),
//   * optionally the fp32 row (exact-fp32 GEMM mode),
//   * the 8 keep bits (1 byte) for the backward.
// With ids == nullptr the kernel is the identity gather (row m of `table`): the user encoder
// uses it to turn its fp32 input [n_users*H, D] into an image.
#pragma once
#include "common.cuh"
#include "gemm_img.cuh"

namespace nrms {

struct GatherArgs {
    const float* table;
    const int64_t* ids;   // [M] or nullptr (identity)
    long long M;          // token rows
    long long vocab;      // rows of `table`
    int D;
    float* x_f32;         // optional [M, D]
    ig::Img x_img;        // optional (hi == nullptr: skip)
    uint8_t* mask;        // optional keep bits [M, mask_bytes]
    int mask_bytes;
    Dropout drop;         // stream kDropEmbedding
};

__device__ __forceinline__ void gather_emit(const GatherArgs& a, bool img, long long m, int g, float4 v0, float4 v1) {
    const int c = g * 8;
    float x[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
    if (a.drop.enabled() && c < a.D) {
        const uint32_t keep = a.drop.keep8(kDropEmbedding, (uint64_t)m, (uint32_t)g);
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = ((keep >> j) & 1u) ? x[j] * a.drop.scale : 0.f;
        if (a.mask && g < a.mask_bytes) a.mask[m * a.mask_bytes + g] = (uint8_t)keep;
    }
    // image column D carries 1.0: the weight-gradient GEMM then yields the bias gradient as
    // its column D (sum_t dY[t,j] * 1); the forward GEMM's weight image is zero there
    if (img && c <= a.D && a.D < c + 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (c + j == a.D) x[j] = 1.f;
    }
    if (a.x_f32) {
        if (c < a.D) *reinterpret_cast<float4*>(a.x_f32 + m * a.D + c) = make_float4(x[0], x[1], x[2], x[3]);
        if (c + 4 < a.D) *reinterpret_cast<float4*>(a.x_f32 + m * a.D + c + 4) = make_float4(x[4], x[5], x[6], x[7]);
    }
    if (img) ig::img_store8(a.x_img, m, g, x);
}

// Work items are (row, 8-column group) pairs, flattened: with D = 300 a row has 40 groups (5 image
// chunks), so a warp that walked one row at a time would run a second, quarter-full pass per row;
// flattened, four rows are exactly five full passes.  A thread carries TWO items per pass so that
// the dependent load chains (token id -> table row) of the two overlap (K = items in flight per thread;
// K = 4 measured slower on B200: 64 -> 72 us at cfg2, 80 registers).
template <int K>
__global__ void __launch_bounds__(256) gather_rows_img_kernel(const GatherArgs a) {
    const bool img = a.x_img.hi != nullptr;
    const long long rows = img ? a.x_img.rows_pad : a.M;
    const int groups = img ? a.x_img.chunks * 8 : ceil_div(a.D, 8);
    const long long total = rows * groups;
    const long long nthreads = (long long)gridDim.x * blockDim.x;
    for (long long it0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; it0 < total; it0 += K * nthreads) {
        long long m[K];
        int g[K];
        bool live[K], real[K];
        float4 v0[K], v1[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const long long it = it0 + k * nthreads;
            live[k] = it < total;
            m[k] = live[k] ? it / groups : 0;
            g[k] = live[k] ? (int)(it - m[k] * groups) : 0;
            real[k] = live[k] && m[k] < a.M;
        }
        long long src[K];
#pragma unroll
        for (int k = 0; k < K; ++k) src[k] = real[k] ? (a.ids ? __ldg(a.ids + m[k]) : m[k]) : -1;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const bool ok = real[k] && src[k] >= 0 && src[k] < a.vocab;
            const float* row = a.table + (ok ? src[k] : 0) * a.D;
            const int c = g[k] * 8;
            v0[k] = (ok && c < a.D) ? __ldg(reinterpret_cast<const float4*>(row + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
            v1[k] = (ok && c + 4 < a.D) ? __ldg(reinterpret_cast<const float4*>(row + c + 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int k = 0; k < K; ++k) {
            if (!live[k]) continue;
            if (!real[k]) {   // image pad rows: zero (they enter the weight-gradient reduction)
                ig::img_store8_zero(a.x_img, m[k], g[k]);
                continue;
            }
            gather_emit(a, img, m[k], g[k], v0[k], v1[k]);
        }
    }
}

}  // namespace nrms
