// pooling.cuh — additive-attention pooling (AdditiveAttention.forward nrms_v0.py:100-126)
// after the tanh(linear) projection, and its backward.
//   a_l = t_l . q            (t = tanh(ctx W_a^T + b_a) comes from the projection GEMM)
//   w   = softmax_l(a)       (dim=1: over the sequence; pad tokens participate)
//   out = sum_l w_l ctx_l
#pragma once
#include "common.cuh"
#include "gemm_img.cuh"

namespace nrms {

struct PoolArgs {
    const float* ctx;    // [M, D], or nullptr: read the context from its split-bf16 image instead
    ig::Img ctx_img;     //   (x = hi + lo, 2^-17 relative: the values the projection GEMM sees)
    const float* t;      // [M, Q]
    const float* q;      // [Q]
    const float* score;  // [M] a_l = t_l . q when the projection GEMM's epilogue already made it, else nullptr
    float* w;            // [n_seq, L]  fwd out / bwd in
    float* out;          // [n_seq, D]  fwd out
    const float* d_out;  // [n_seq, D]  bwd in
    float* d_ctx;        // [M, D]      bwd out (optional): w_l * d_out; the projection path is added by the
                         //             data-gradient GEMM, whose tcgen05 epilogue forms this term itself
    float* d_pre;        // [M, Q]      bwd out (optional): grad wrt pre-tanh activations, fp32
    ig::Img d_pre_img;   //             bwd out (optional): the same as a split-bf16 image
    float* d_part;       // [n_seq, 2Q] bwd out: per-sequence partials of (d_b_a | d_q)
    long long M;
    int L, D, Q;
};

// columns [8u, 8u+8) of row r of a split-bf16 image as fp32 (hi + lo; hi alone for a single-plane image)
__device__ __forceinline__ void img_load8(const ig::Img& im, long long r, int u, float* x) {
    const long long off = ig::img_unit_off(im.chunk_stride, r, u);
    const uint4 h = __ldg(reinterpret_cast<const uint4*>(im.hi + off));
    const uint4 l = im.lo ? __ldg(reinterpret_cast<const uint4*>(im.lo + off)) : make_uint4(0u, 0u, 0u, 0u);
    const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        x[2 * j] = __uint_as_float(hw[j] << 16) + __uint_as_float(lw[j] << 16);
        x[2 * j + 1] = __uint_as_float(hw[j] & 0xffff0000u) + __uint_as_float(lw[j] & 0xffff0000u);
    }
}

// one CTA per sequence; dynamic smem: L floats (+ 8 * ceil(D/8) * (256 / ceil(D/8)) with an image context)
__global__ void __launch_bounds__(256) pool_fwd_kernel(const PoolArgs p) {
    extern __shared__ float sw[];
    const int seq = blockIdx.x, L = p.L, D = p.D, Q = p.Q;
    const long long row0 = (long long)seq * L;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    if (p.score) {
        for (int l = threadIdx.x; l < L; l += blockDim.x) sw[l] = p.score[row0 + l];
    } else {
        for (int l = warp; l < L; l += nw) {
            const float* tr = p.t + (row0 + l) * Q;
            float s = 0.f;
            for (int j = lane; j < Q; j += 32) s = fmaf(__ldg(tr + j), __ldg(p.q + j), s);
            s = warp_sum(s);
            if (lane == 0) sw[l] = s;
        }
    }
    __syncthreads();
    if (warp == 0) {
        float mx = -INFINITY;
        for (int l = lane; l < L; l += 32) mx = fmaxf(mx, sw[l]);
        mx = warp_max(mx);
        float den = 0.f;
        for (int l = lane; l < L; l += 32) {
            const float e = __expf(sw[l] - mx);
            sw[l] = e;
            den += e;
        }
        den = warp_sum(den);
        const float inv = 1.f / den;
        for (int l = lane; l < L; l += 32) {
            const float wv = sw[l] * inv;
            sw[l] = wv;
            if (p.w) p.w[(long long)seq * L + l] = wv;
        }
    }
    __syncthreads();
    if (p.ctx == nullptr) {
        // image context: a thread owns one 8-column unit for every `slices`-th row, then the slices are summed
        const int units = ceil_div(D, 8), slices = blockDim.x / units;
        float* part = sw + L;                              // [slices][units * 8]
        const int u = threadIdx.x % units, sl = threadIdx.x / units;
        if (sl < slices) {
            float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            for (int l = sl; l < L; l += slices) {
                float x[8];
                img_load8(p.ctx_img, row0 + l, u, x);
                const float wl = sw[l];
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] = fmaf(wl, x[j], acc[j]);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) part[(sl * units + u) * 8 + j] = acc[j];
        }
        __syncthreads();
        for (int d = threadIdx.x; d < D; d += blockDim.x) {
            float acc = 0.f;
            for (int s2 = 0; s2 < slices; ++s2) acc += part[s2 * units * 8 + d];
            p.out[(long long)seq * D + d] = acc;
        }
        return;
    }
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        float acc0 = 0.f, acc1 = 0.f;
        int l = 0;
        for (; l + 1 < L; l += 2) {
            acc0 = fmaf(sw[l], __ldg(p.ctx + (row0 + l) * D + d), acc0);
            acc1 = fmaf(sw[l + 1], __ldg(p.ctx + (row0 + l + 1) * D + d), acc1);
        }
        if (l < L) acc0 = fmaf(sw[l], __ldg(p.ctx + (row0 + l) * D + d), acc0);
        p.out[(long long)seq * D + d] = acc0 + acc1;
    }
}

// one CTA per sequence; dynamic smem: 2L floats (w, da)
__global__ void __launch_bounds__(256) pool_bwd_kernel(const PoolArgs p) {
    extern __shared__ float sm[];
    const int seq = blockIdx.x, L = p.L, D = p.D, Q = p.Q;
    float* sw = sm;
    float* sda = sm + L;
    __shared__ float s_dot;
    const long long row0 = (long long)seq * L;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const float* go = p.d_out + (long long)seq * D;
    // dw_l = d_out . ctx_l
    for (int l = warp; l < L; l += nw) {
        float s = 0.f;
        if (p.ctx == nullptr) {
            for (int u = lane; 8 * u < D; u += 32) {       // image context: a lane owns 8-column units
                float x[8];
                img_load8(p.ctx_img, row0 + l, u, x);
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (8 * u + j < D) s = fmaf(__ldg(go + 8 * u + j), x[j], s);
            }
        } else {
            const float* cr = p.ctx + (row0 + l) * D;
            for (int d = lane; d < D; d += 32) s = fmaf(__ldg(go + d), __ldg(cr + d), s);
        }
        s = warp_sum(s);
        if (lane == 0) {
            sda[l] = s;
            sw[l] = p.w[(long long)seq * L + l];
        }
    }
    __syncthreads();
    if (warp == 0) {
        float s = 0.f;
        for (int l = lane; l < L; l += 32) s = fmaf(sw[l], sda[l], s);
        s = warp_sum(s);
        if (lane == 0) s_dot = s;
    }
    __syncthreads();
    const float dot = s_dot;
    __syncthreads();
    for (int l = threadIdx.x; l < L; l += blockDim.x) sda[l] = sw[l] * (sda[l] - dot);
    __syncthreads();
    // d_ctx (pooling path)
    if (p.d_ctx) {
        const int d4n = D >> 2;
        for (int i = threadIdx.x; i < L * d4n; i += blockDim.x) {
            const int l = i / d4n, d = (i - l * d4n) << 2;
            const float4 g = __ldg(reinterpret_cast<const float4*>(go + d));
            const float wl = sw[l];
            *reinterpret_cast<float4*>(p.d_ctx + (row0 + l) * D + d) = make_float4(wl * g.x, wl * g.y, wl * g.z, wl * g.w);
        }
    }
    // d_pre = da_l * q_j * (1 - t^2), coalesced 2-column units (fp32 and/or image)
    const bool img = p.d_pre_img.hi != nullptr;
    {
        if (img && !p.d_pre && Q % 8 == 0) {
            // image only (the tcgen05 path): a thread owns one 16-byte unit (8 columns) of a row
            const int q8 = Q >> 3;
            for (int i = threadIdx.x; i < L * q8; i += blockDim.x) {
                const int l = i / q8, u = i - l * q8, j = u << 3;
                const float4 t0 = __ldg(reinterpret_cast<const float4*>(p.t + (row0 + l) * Q + j));
                const float4 t1 = __ldg(reinterpret_cast<const float4*>(p.t + (row0 + l) * Q + j + 4));
                const float4 q0 = __ldg(reinterpret_cast<const float4*>(p.q + j));
                const float4 q1 = __ldg(reinterpret_cast<const float4*>(p.q + j + 4));
                const float da = sda[l];
                const float x[8] = {da * q0.x * (1.f - t0.x * t0.x), da * q0.y * (1.f - t0.y * t0.y),
                                    da * q0.z * (1.f - t0.z * t0.z), da * q0.w * (1.f - t0.w * t0.w),
                                    da * q1.x * (1.f - t1.x * t1.x), da * q1.y * (1.f - t1.y * t1.y),
                                    da * q1.z * (1.f - t1.z * t1.z), da * q1.w * (1.f - t1.w * t1.w)};
                ig::img_store8(p.d_pre_img, row0 + l, u, x);
            }
        } else {
        const int q2 = Q >> 1;
        for (int i = threadIdx.x; i < L * q2; i += blockDim.x) {
            const int l = i / q2, j = (i - l * q2) << 1;
            const float2 tv = __ldg(reinterpret_cast<const float2*>(p.t + (row0 + l) * Q + j));
            const float2 qv = __ldg(reinterpret_cast<const float2*>(p.q + j));
            const float da = sda[l];
            const float2 dp = make_float2(da * qv.x * (1.f - tv.x * tv.x), da * qv.y * (1.f - tv.y * tv.y));
            if (p.d_pre) *reinterpret_cast<float2*>(p.d_pre + (row0 + l) * Q + j) = dp;
            if (img) {
                __nv_bfloat16 h0b, l0b, h1b, l1b;
                tc::split_bf16(dp.x, h0b, l0b);
                tc::split_bf16(dp.y, h1b, l1b);
                const long long off = ig::img_unit_off(p.d_pre_img.chunk_stride, row0 + l, j >> 3) + (j & 7) * 2;
                *reinterpret_cast<uint32_t*>(p.d_pre_img.hi + off) =
                    (uint32_t)__bfloat16_as_ushort(h0b) | ((uint32_t)__bfloat16_as_ushort(h1b) << 16);
                if (p.d_pre_img.lo)
                    *reinterpret_cast<uint32_t*>(p.d_pre_img.lo + off) =
                        (uint32_t)__bfloat16_as_ushort(l0b) | ((uint32_t)__bfloat16_as_ushort(l1b) << 16);
            }
        }
        }
        if (img) {
            // zero padding: columns [Q, 16*ceil(Q/16)) feed the data-gradient GEMM's last k-step,
            // rows [M, rows_pad) the weight-gradient reduction
            const int cpad = ceil_div(Q, 16) * 16 - Q;   // even (Q % 4 == 0)
            for (int i = threadIdx.x; i < L * (cpad >> 1); i += blockDim.x) {
                const int l = i / (cpad >> 1), j = Q + ((i - l * (cpad >> 1)) << 1);
                const long long off = ig::img_unit_off(p.d_pre_img.chunk_stride, row0 + l, j >> 3) + (j & 7) * 2;
                *reinterpret_cast<uint32_t*>(p.d_pre_img.hi + off) = 0u;
                if (p.d_pre_img.lo) *reinterpret_cast<uint32_t*>(p.d_pre_img.lo + off) = 0u;
            }
            if (seq == (int)gridDim.x - 1) {
                const int groups = p.d_pre_img.chunks * 8;
                const long long npad = p.d_pre_img.rows_pad - p.M;
                for (long long i = threadIdx.x; i < npad * groups; i += blockDim.x)
                    ig::img_store8_zero(p.d_pre_img, p.M + i / groups, (int)(i % groups));
            }
        }
    }
    // bias / query-vector partials (fixed summation order over l)
    for (int j = threadIdx.x; j < Q; j += blockDim.x) {
        const float qj = __ldg(p.q + j);
        float db = 0.f, dq = 0.f;
        for (int l = 0; l < L; ++l) {
            const float tv = __ldg(p.t + (row0 + l) * Q + j);
            const float da = sda[l];
            db += da * qj * (1.f - tv * tv);
            dq = fmaf(da, tv, dq);
        }
        p.d_part[(long long)seq * 2 * Q + j] = db;
        p.d_part[(long long)seq * 2 * Q + Q + j] = dq;
    }
}

// ---------------------------------------------------------------------------------------------------------
// Second generation of the two kernels for the tensor-core path (image context, scores from the projection
// GEMM's epilogue, image-only d_pre), for launches of FEW CTAs (the user encoder: one CTA per impression).
// The kernels above walk a sequence's rows in DEPENDENT rounds — a warp per row with a shuffle reduction per
// round, then a per-column loop over all rows for the partials — which is what a launch that cannot fill the
// GPU pays in full.  Here every phase issues ALL of its loads before it consumes the first one (work items
// flattened over (row, 16-byte unit), R items per thread in flight), row sums go through a shared [L][units]
// table in a fixed order, and the bias / query-vector partials are accumulated by the threads that form d_pre
// (t is read once, not twice).  They need ~80 registers (3 CTAs per SM instead of 8): measured slower for the
// title encoder's thousands of CTAs, which are throughput-bound (abi.cu: kPool2MaxSeq).
// Same results as the kernels above up to the summation order (fixed, run-to-run identical).
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void unpack8(const uint4& h, const uint4& l, float* x) {
    const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        x[2 * j] = __uint_as_float(hw[j] << 16) + __uint_as_float(lw[j] << 16);
        x[2 * j + 1] = __uint_as_float(hw[j] & 0xffff0000u) + __uint_as_float(lw[j] & 0xffff0000u);
    }
}

constexpr int kPool2R = 5;      // (row, unit) items a thread keeps in flight

// dynamic smem: L floats + 8 * units * (256 / units) floats, units = ceil(D / 8)   (as pool_fwd_kernel)
__global__ void __launch_bounds__(256, 3) pool_fwd2_kernel(const PoolArgs p) {
    extern __shared__ float sw[];
    const int seq = blockIdx.x, L = p.L, D = p.D;
    const long long row0 = (long long)seq * L;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int units = ceil_div(D, 8), slices = blockDim.x / units;
    const int u = threadIdx.x % units, sl = threadIdx.x / units;
    const bool active = sl < slices, has_lo = p.ctx_img.lo != nullptr;
    const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
    uint4 xh[kPool2R], xl[kPool2R];
    // the first batch of context rows leaves for the registers before the scores are even read
#pragma unroll
    for (int k = 0; k < kPool2R; ++k) {
        const int l = sl + k * slices;
        xh[k] = xl[k] = z4;
        if (active && l < L) {
            const long long off = ig::img_unit_off(p.ctx_img.chunk_stride, row0 + l, u);
            xh[k] = __ldg(reinterpret_cast<const uint4*>(p.ctx_img.hi + off));
            if (has_lo) xl[k] = __ldg(reinterpret_cast<const uint4*>(p.ctx_img.lo + off));
        }
    }
    for (int l = threadIdx.x; l < L; l += blockDim.x) sw[l] = p.score[row0 + l];
    __syncthreads();
    // softmax statistics: every warp computes them itself (the same operations in the same order, so all
    // warps hold the same bits) — no second barrier, no idle warps
    float mx = -INFINITY;
    for (int l = lane; l < L; l += 32) mx = fmaxf(mx, sw[l]);
    mx = warp_max(mx);
    float den = 0.f;
    for (int l = lane; l < L; l += 32) den += __expf(sw[l] - mx);
    den = warp_sum(den);
    const float inv = 1.f / den;
    if (warp == 0 && p.w)
        for (int l = lane; l < L; l += 32) p.w[(long long)seq * L + l] = __expf(sw[l] - mx) * inv;
    float* part = sw + L;                                  // [slices][units * 8]
    if (active) {
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int l0 = sl; l0 < L; l0 += kPool2R * slices) {
            if (l0 != sl) {                                // later batches (L > kPool2R * slices)
#pragma unroll
                for (int k = 0; k < kPool2R; ++k) {
                    const int l = l0 + k * slices;
                    xh[k] = xl[k] = z4;
                    if (l < L) {
                        const long long off = ig::img_unit_off(p.ctx_img.chunk_stride, row0 + l, u);
                        xh[k] = __ldg(reinterpret_cast<const uint4*>(p.ctx_img.hi + off));
                        if (has_lo) xl[k] = __ldg(reinterpret_cast<const uint4*>(p.ctx_img.lo + off));
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < kPool2R; ++k) {
                const int l = l0 + k * slices;
                if (l < L) {
                    float x[8];
                    unpack8(xh[k], xl[k], x);
                    const float wl = __expf(sw[l] - mx) * inv;
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[j] = fmaf(wl, x[j], acc[j]);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) part[(sl * units + u) * 8 + j] = acc[j];
    }
    __syncthreads();
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        float acc = 0.f;
        for (int s2 = 0; s2 < slices; ++s2) acc += part[s2 * units * 8 + d];
        p.out[(long long)seq * D + d] = acc;
    }
}

// floats of dynamic shared memory of pool_bwd2_kernel
__host__ __device__ inline int pool_bwd2_smem_floats(int L, int D, int Q) {
    const int a = L * ceil_div(D, 8), b = (256 / (Q >> 3)) * 2 * Q;
    return ((2 * L + 3) & ~3) + 8 * ceil_div(D, 8) + (a > b ? a : b);
}

// needs: image context, Q % 8 == 0, 8 <= Q <= 2048, image-only d_pre (no fp32 d_pre / d_ctx outputs)
__global__ void __launch_bounds__(256, 3) pool_bwd2_kernel(const PoolArgs p) {
    extern __shared__ float sm[];
    const int seq = blockIdx.x, L = p.L, D = p.D, Q = p.Q;
    float* sw = sm;
    float* sda = sm + L;
    const int units = ceil_div(D, 8), total = L * units;
    float* sgo = sm + ((2 * L + 3) & ~3);                  // d_out of the sequence (16-byte aligned), zero beyond column D
    float* spart = sgo + 8 * units;                        // [L][units] dot pieces, later [slices][2Q] partials
    __shared__ float s_dot;
    const long long row0 = (long long)seq * L;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const float* go = p.d_out + (long long)seq * D;
    const bool has_lo = p.ctx_img.lo != nullptr;
    const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
    const float4 f0 = make_float4(0.f, 0.f, 0.f, 0.f);
    // ---- dw_l = d_out . ctx_l: one (row, unit) item = 8 columns; all of a thread's items in flight at once
    uint4 xh[kPool2R], xl[kPool2R];
#pragma unroll
    for (int k = 0; k < kPool2R; ++k) {                    // first batch: issued before anything else
        const int i = (int)threadIdx.x + k * (int)blockDim.x;
        xh[k] = xl[k] = z4;
        if (i < total) {
            const int l = i / units, u = i - l * units;
            const long long off = ig::img_unit_off(p.ctx_img.chunk_stride, row0 + l, u);
            xh[k] = __ldg(reinterpret_cast<const uint4*>(p.ctx_img.hi + off));
            if (has_lo) xl[k] = __ldg(reinterpret_cast<const uint4*>(p.ctx_img.lo + off));
        }
    }
    for (int l = threadIdx.x; l < L; l += blockDim.x) sw[l] = p.w[(long long)seq * L + l];
    // columns >= D of the image are not context (the ones column of the weight-gradient GEMM, padding): zero weight
    for (int d = threadIdx.x; d < 8 * units; d += blockDim.x) sgo[d] = d < D ? __ldg(go + d) : 0.f;
    __syncthreads();
    for (int i0 = threadIdx.x; i0 < total; i0 += kPool2R * (int)blockDim.x) {
        if (i0 != (int)threadIdx.x) {                      // later batches (L * units > kPool2R * 256)
#pragma unroll
            for (int k = 0; k < kPool2R; ++k) {
                const int i = i0 + k * (int)blockDim.x;
                xh[k] = xl[k] = z4;
                if (i < total) {
                    const int l = i / units, u = i - l * units;
                    const long long off = ig::img_unit_off(p.ctx_img.chunk_stride, row0 + l, u);
                    xh[k] = __ldg(reinterpret_cast<const uint4*>(p.ctx_img.hi + off));
                    if (has_lo) xl[k] = __ldg(reinterpret_cast<const uint4*>(p.ctx_img.lo + off));
                }
            }
        }
#pragma unroll
        for (int k = 0; k < kPool2R; ++k) {
            const int i = i0 + k * (int)blockDim.x;
            if (i < total) {
                const int u = i % units;
                float x[8];
                unpack8(xh[k], xl[k], x);
                const float4 g0 = *reinterpret_cast<const float4*>(sgo + 8 * u);
                const float4 g1 = *reinterpret_cast<const float4*>(sgo + 8 * u + 4);
                float s = g0.x * x[0];
                s = fmaf(g0.y, x[1], s);
                s = fmaf(g0.z, x[2], s);
                s = fmaf(g0.w, x[3], s);
                s = fmaf(g1.x, x[4], s);
                s = fmaf(g1.y, x[5], s);
                s = fmaf(g1.z, x[6], s);
                s = fmaf(g1.w, x[7], s);
                spart[i] = s;
            }
        }
    }
    __syncthreads();
    for (int l = warp; l < L; l += nw) {
        float s = 0.f;
        for (int u = lane; u < units; u += 32) s += spart[l * units + u];
        s = warp_sum(s);
        if (lane == 0) sda[l] = s;
    }
    __syncthreads();
    if (warp == 0) {
        float s = 0.f;
        for (int l = lane; l < L; l += 32) s = fmaf(sw[l], sda[l], s);
        s = warp_sum(s);
        if (lane == 0) s_dot = s;
    }
    __syncthreads();
    const float dot = s_dot;
    for (int l = threadIdx.x; l < L; l += blockDim.x) sda[l] = sw[l] * (sda[l] - dot);
    __syncthreads();                                       // (also: every read of spart's dot pieces is over)
    // ---- d_pre = da_l * q_j * (1 - t^2) as 16-byte image units, and the per-sequence partials of
    //      d_b_a = sum_l d_pre[l, j] and d_q = sum_l da_l * t[l, j] from the same registers
    const int q8 = Q >> 3, slices = (int)blockDim.x / q8;
    {
        const int u = threadIdx.x % q8, sl = threadIdx.x / q8;
        if (sl < slices) {
            const float4 q0 = __ldg(reinterpret_cast<const float4*>(p.q + 8 * u));
            const float4 q1 = __ldg(reinterpret_cast<const float4*>(p.q + 8 * u + 4));
            const float qv[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
            float db[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, dq[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            constexpr int R3 = 4;
            for (int l0 = sl; l0 < L; l0 += R3 * slices) {
                float4 t0[R3], t1[R3];
#pragma unroll
                for (int k = 0; k < R3; ++k) {
                    const int l = l0 + k * slices;
                    t0[k] = t1[k] = f0;
                    if (l < L) {
                        t0[k] = __ldg(reinterpret_cast<const float4*>(p.t + (row0 + l) * Q + 8 * u));
                        t1[k] = __ldg(reinterpret_cast<const float4*>(p.t + (row0 + l) * Q + 8 * u + 4));
                    }
                }
#pragma unroll
                for (int k = 0; k < R3; ++k) {
                    const int l = l0 + k * slices;
                    if (l < L) {
                        const float tv[8] = {t0[k].x, t0[k].y, t0[k].z, t0[k].w, t1[k].x, t1[k].y, t1[k].z, t1[k].w};
                        const float da = sda[l];
                        float x[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            x[j] = da * qv[j] * (1.f - tv[j] * tv[j]);
                            db[j] += x[j];
                            dq[j] = fmaf(da, tv[j], dq[j]);
                        }
                        ig::img_store8(p.d_pre_img, row0 + l, u, x);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                spart[sl * 2 * Q + 8 * u + j] = db[j];
                spart[sl * 2 * Q + Q + 8 * u + j] = dq[j];
            }
        }
    }
    // zero padding of the d_pre image (as pool_bwd_kernel): columns [Q, 16*ceil(Q/16)) feed the data-gradient
    // GEMM's last k-step, rows [M, rows_pad) the weight-gradient reduction
    {
        const int cpad = ceil_div(Q, 16) * 16 - Q;         // 0 or 8 (Q % 8 == 0)
        for (int i = threadIdx.x; i < L * (cpad >> 1); i += blockDim.x) {
            const int l = i / (cpad >> 1), j = Q + ((i - l * (cpad >> 1)) << 1);
            const long long off = ig::img_unit_off(p.d_pre_img.chunk_stride, row0 + l, j >> 3) + (j & 7) * 2;
            *reinterpret_cast<uint32_t*>(p.d_pre_img.hi + off) = 0u;
            if (p.d_pre_img.lo) *reinterpret_cast<uint32_t*>(p.d_pre_img.lo + off) = 0u;
        }
        if (seq == (int)gridDim.x - 1) {
            const int groups = p.d_pre_img.chunks * 8;
            const long long npad = p.d_pre_img.rows_pad - p.M;
            for (long long i = threadIdx.x; i < npad * groups; i += blockDim.x)
                ig::img_store8_zero(p.d_pre_img, p.M + i / groups, (int)(i % groups));
        }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < 2 * Q; j += blockDim.x) {
        float acc = 0.f;
        for (int s2 = 0; s2 < slices; ++s2) acc += spart[s2 * 2 * Q + j];
        p.d_part[(long long)seq * 2 * Q + j] = acc;
    }
}

// out[n] (+)= scale * sum_{r<R} in[r, n]   (deterministic two-level: each thread walks a
// column; R is at most a few thousand).  Used for bias / query partials and split-K weight
// gradient partials.
__global__ void reduce_rows_kernel(const float* __restrict__ in, float* __restrict__ out,
                                   long long R, long long n, long long ld, float scale,
                                   int accumulate) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    long long r = 0;
    for (; r + 3 < R; r += 4) {
        s0 += in[(r + 0) * ld + c];
        s1 += in[(r + 1) * ld + c];
        s2 += in[(r + 2) * ld + c];
        s3 += in[(r + 3) * ld + c];
    }
    for (; r < R; ++r) s0 += in[r * ld + c];
    const float s = ((s0 + s1) + (s2 + s3)) * scale;
    out[c] = accumulate ? out[c] + s : s;
}

// Sum of split-K weight-gradient partials whose column `cols` carries the bias gradient (the
// activation image has a ones column, gather.cuh): part [splits][rows][ldp] ->
// dW[r, c] (c < cols, contiguous [rows, cols]) and db[r].  Fixed summation order over the splits.
// hp_dk > 0: partial row r is row ig::hp_unpad(r, hp_D, hp_dk) of dW / db (padding rows skipped).
__global__ void reduce_wgrad_kernel(const float* __restrict__ part, int splits, int rows, int ldp, int cols,
                                    float* __restrict__ dW, float* __restrict__ db, int hp_D, int hp_dk) {
    // a thread owns 4 consecutive columns of a row (ldp and cols are multiples of 4: 16-byte loads,
    // four independent sums, two splits in flight) — the walk over the splits is latency-bound
    const long long per4 = (long long)rows * ldp / 4;
    const int q4 = cols / 4 + 1;                        // float4 units per row: cols/4 of dW, then {db, pad}
    const float4* part4 = reinterpret_cast<const float4*>(part);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)rows * q4;
         i += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(i / q4), c4 = (int)(i - (long long)r * q4);
        const int ro = hp_dk > 0 ? ig::hp_unpad(r, hp_D, hp_dk) : r;
        if (ro < 0) continue;
        const float4* p = part4 + ((long long)r * ldp) / 4 + c4;
        // eight splits in flight per round (the walk over the splits is a chain of L2 round trips: with two in
        // flight the 74 splits of the additive projection's gradient cost 37 of them), combined in a fixed tree
        float4 a[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        int s = 0;
        for (; s + 7 < splits; s += 8) {
            float4 x[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) x[k] = __ldg(p + (s + k) * per4);
#pragma unroll
            for (int k = 0; k < 8; ++k) { a[k].x += x[k].x; a[k].y += x[k].y; a[k].z += x[k].z; a[k].w += x[k].w; }
        }
        {
            float4 x[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) x[k] = s + k < splits ? __ldg(p + (s + k) * per4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int k = 0; k < 8; ++k) { a[k].x += x[k].x; a[k].y += x[k].y; a[k].z += x[k].z; a[k].w += x[k].w; }
        }
#pragma unroll
        for (int w = 4; w > 0; w >>= 1)
#pragma unroll
            for (int k = 0; k < w; ++k) { a[k].x += a[k + w].x; a[k].y += a[k + w].y; a[k].z += a[k + w].z; a[k].w += a[k + w].w; }
        const float4 v = a[0];
        if (4 * c4 < cols) *reinterpret_cast<float4*>(dW + (long long)ro * cols + 4 * c4) = v;
        else db[ro] = v.x;
    }
}

// Column sums with the rows spread over the warps of a CTA: block (32, kRedWarps), warp y of slice
// blockIdx.y takes rows r0 + y, r0 + y + kRedWarps, ... (four loads in flight), the warps' partials are added
// in warp order through shared memory:  out[slice, c] = scale * sum_{r in slice} in[r, c].  One thread per
// column walking all rows (the kernels below) is a chain of R/4 dependent L2 round trips.
constexpr int kRedWarps = 16;
__global__ void __launch_bounds__(32 * kRedWarps) reduce_rows_w_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                                     long long R, long long n, long long ld, long long per,
                                                                     float scale) {
    __shared__ float sh[kRedWarps][33];
    const long long c = (long long)blockIdx.x * 32 + threadIdx.x;
    const long long r0 = (long long)blockIdx.y * per;
    const long long r1 = r0 + per < R ? r0 + per : R;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    if (c < n) {
        long long r = r0 + threadIdx.y;
        for (; r + 3 * kRedWarps < r1; r += 4 * kRedWarps) {
            s0 += in[(r + 0 * kRedWarps) * ld + c];
            s1 += in[(r + 1 * kRedWarps) * ld + c];
            s2 += in[(r + 2 * kRedWarps) * ld + c];
            s3 += in[(r + 3 * kRedWarps) * ld + c];
        }
        for (; r < r1; r += kRedWarps) s0 += in[r * ld + c];
    }
    sh[threadIdx.y][threadIdx.x] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (threadIdx.y == 0 && c < n) {
        float acc = 0.f;
#pragma unroll
        for (int y = 0; y < kRedWarps; ++y) acc += sh[y][threadIdx.x];
        out[(long long)blockIdx.y * n + c] = acc * scale;
    }
}

// first level of the two-level column sum: slice y sums rows [y*per, (y+1)*per) into tmp[y, :]
__global__ void reduce_rows_sliced_kernel(const float* __restrict__ in, float* __restrict__ tmp,
                                          long long R, long long n, long long ld, long long per) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    const long long r0 = (long long)blockIdx.y * per;
    const long long r1 = r0 + per < R ? r0 + per : R;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    long long r = r0;
    for (; r + 3 < r1; r += 4) {
        s0 += in[(r + 0) * ld + c];
        s1 += in[(r + 1) * ld + c];
        s2 += in[(r + 2) * ld + c];
        s3 += in[(r + 3) * ld + c];
    }
    for (; r < r1; ++r) s0 += in[r * ld + c];
    tmp[(long long)blockIdx.y * n + c] = (s0 + s1) + (s2 + s3);
}

}  // namespace nrms
