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
        float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;
        int s = 0;
        for (; s + 1 < splits; s += 2) {
            const float4 x = __ldg(p + s * per4), y = __ldg(p + (s + 1) * per4);
            s0.x += x.x; s0.y += x.y; s0.z += x.z; s0.w += x.w;
            s1.x += y.x; s1.y += y.y; s1.z += y.z; s1.w += y.w;
        }
        if (s < splits) {
            const float4 x = __ldg(p + s * per4);
            s0.x += x.x; s0.y += x.y; s0.z += x.z; s0.w += x.w;
        }
        const float4 v = make_float4(s0.x + s1.x, s0.y + s1.y, s0.z + s1.z, s0.w + s1.w);
        if (4 * c4 < cols) *reinterpret_cast<float4*>(dW + (long long)ro * cols + 4 * c4) = v;
        else db[ro] = v.x;
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
