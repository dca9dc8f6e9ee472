// masked_user.cuh — kernels of the `nrms` sibling variant (reference model/nrms.py).
//
// That model's news vectors are rows of a BERT-vector table passed through one Linear; its user encoder is
// an 8-head self-attention over the 50-slot history with a PADDING MASK, dropout on the attention
// PROBABILITIES and an output projection, followed by an additive attention with the same mask
// (nrms.py:26-117,258-271).  Head dims are 64-100 and the model dim 512: outside the head-padded
// 32-column tiling of attention_hp*.cuh, and tiny next to the title encoder this library is built around
// (B x 8 items of 50 x 50 x 64 per step), so the attention and pooling here are fp32 CUDA-core kernels —
// one CTA per (sequence, head) / per sequence, operands in shared memory, a warp per row — while the four
// Linears run on the tcgen05 image GEMMs (abi_masked.inc).  Every reduction runs in a fixed order.
#pragma once
#include "common.cuh"
#include "profiler.cuh"

namespace nrms {
namespace mu {

constexpr uint32_t kDropCandVec = 3u;   // nrms.py:254 on the candidate vectors (:339)
constexpr uint32_t kDropHistVec = 4u;   // nrms.py:254 on the history vectors (:343)
constexpr uint32_t kDropAttnProb = 5u;  // nrms.py:45-47 on softmax(QK^T) of the user encoder

// ---- y = x * keep / (1 - p), Philox-addressed like every dropout of the library (common.cuh) -----------
__global__ void dropout_apply_kernel(Dropout dr, uint32_t sid, long long n_rows, int n_cols,
                                     const float* __restrict__ x, float* __restrict__ y) {
    const int groups = ceil_div(n_cols, 8);
    const long long total = n_rows * groups;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / groups;
        const int g = (int)(i - r * groups);
        const uint32_t keep = dr.enabled() ? dr.keep8(sid, (uint64_t)r, (uint32_t)g) : 0xffu;
        for (int j = 0; j < 8 && g * 8 + j < n_cols; ++j) {
            const long long o = r * n_cols + g * 8 + j;
            y[o] = ((keep >> j) & 1u) ? x[o] * dr.scale : 0.f;
        }
    }
}

// ---- masked multi-head attention (nrms.py:26-49) ---------------------------------------------------------
// qkv [B, L, 3E] fp32 (Q | K | V of every head side by side, the fused projection's output), mask [B, L]
// (1 = real slot) or NULL, probs [B, h, L, L] = softmax BEFORE dropout (saved for the backward),
// ctx [B, L, E].  One CTA per (sequence, head), 8 warps, every operand of the item in shared memory as a
// row-major fp32 tile whose row length ld is a multiple of 4 with ld/4 odd (16-byte loads of 8 different
// rows hit 8 different bank groups).  Register tiling: a warp owns EIGHT query rows at a time —
//   rows x keys  (S = Q K^T, dPd = dO V^T): lane = key (j = lane + 32 t), k = head columns, 4 at a time:
//                 8 broadcast 16-byte loads of the rows + T 16-byte loads of the lane's keys per 32 T FMAs;
//   rows x cols  (O = P V, dQ = dS K, and with keys for rows dK = dS^T Q, dV = Pd^T dO): lane = head column,
//                 k = keys; the 8 weights of a key are two broadcast 16-byte loads (the warp stages its P /
//                 dS rows transposed, [key][8]; dK / dV read 8 adjacent columns of the row-major dS / Pd).
// That is ~4 FMAs per shared-memory wavefront (the first version, one row per warp, had 0.5).
struct MaskedAttnArgs {
    const float* qkv;
    const uint8_t* mask;
    float* probs;
    float* ctx;
    const float* d_ctx;   // backward
    float* d_qkv;         // backward: [B, L, 3E]
    int B, L, heads, dk;
    float scale;          // 1 / sqrt(dk)
    Dropout drop;
};
constexpr int kMaWarps = 8;
constexpr int kMaRows = 8;                // rows per warp step
constexpr int kMaMaxKeysPerLane = 4;      // L <= 128
constexpr int kMaMaxColsPerLane = 4;      // head dim <= 128

__host__ __device__ inline int ma_ld(int dk) {
    int ld = (dk + 3) / 4 * 4;
    if (((ld / 4) & 1) == 0) ld += 4;
    return ld;
}
__host__ __device__ inline int ma_lp(int L) { return (L + 7) / 8 * 8; }

inline size_t masked_attn_fwd_smem(int L, int dk) {
    return sizeof(float) * ((size_t)3 * L * ma_ld(dk) + (size_t)kMaWarps * L * kMaRows);
}
inline size_t masked_attn_bwd_smem(int L, int dk) {
    return sizeof(float) * ((size_t)4 * L * ma_ld(dk) + (size_t)2 * L * ma_lp(L) + (size_t)kMaWarps * L * kMaRows);
}

// rows [0, L) x columns [0, dk) of a [.., stride] global matrix -> tile [L][ld], padding columns zero.
// 16-byte cp.async where the source allows it (every load of the CTA in flight at once: the tiles are the
// only global reads of the kernel and with 2 CTAs per SM their latency is exposed); the caller waits with
// ma_tiles_wait() before its __syncthreads().
__device__ __forceinline__ void ma_load_tile(float* dst, const float* __restrict__ src, long long stride, int L, int dk, int ld) {
    const bool vec = (dk & 3) == 0 && (stride & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0;
    if (vec) {
        const int c4 = dk >> 2;
        for (int i = threadIdx.x; i < L * c4; i += blockDim.x) {
            const int r = i / c4, c = i - r * c4;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst + r * ld + 4 * c)),
                         "l"(src + (long long)r * stride + 4 * c) : "memory");
        }
        for (int i = threadIdx.x; i < L * (ld - dk); i += blockDim.x) {
            const int r = i / (ld - dk), d = dk + i % (ld - dk);
            dst[r * ld + d] = 0.f;
        }
    } else {
        for (int i = threadIdx.x; i < L * ld; i += blockDim.x) {
            const int r = i / ld, d = i - r * ld;
            dst[i] = d < dk ? src[(long long)r * stride + d] : 0.f;
        }
    }
}
__device__ __forceinline__ void ma_tiles_wait() { asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory"); }
// acc[r][t] = sum_d A[i0 + r][d] * Bm[lane + 32 t][d]   (rows / keys beyond L read row L - 1: discarded by the caller)
template <int T>
__device__ __forceinline__ void ma_rows_dot(float (&acc)[kMaRows][T], const float* A, const float* Bm, int i0, int L, int ld,
                                            int lane) {
    const float* brow[T];
    const float* arow[kMaRows];
#pragma unroll
    for (int t = 0; t < T; ++t) brow[t] = Bm + min(lane + 32 * t, L - 1) * ld;
#pragma unroll
    for (int r = 0; r < kMaRows; ++r) {
        arow[r] = A + min(i0 + r, L - 1) * ld;
#pragma unroll
        for (int t = 0; t < T; ++t) acc[r][t] = 0.f;
    }
    for (int d = 0; d < ld; d += 4) {
        float4 bv[T];
#pragma unroll
        for (int t = 0; t < T; ++t) bv[t] = *reinterpret_cast<const float4*>(brow[t] + d);
#pragma unroll
        for (int r = 0; r < kMaRows; ++r) {
            const float4 av = *reinterpret_cast<const float4*>(arow[r] + d);
#pragma unroll
            for (int t = 0; t < T; ++t) {
                acc[r][t] = fmaf(av.x, bv[t].x, acc[r][t]);
                acc[r][t] = fmaf(av.y, bv[t].y, acc[r][t]);
                acc[r][t] = fmaf(av.z, bv[t].z, acc[r][t]);
                acc[r][t] = fmaf(av.w, bv[t].w, acc[r][t]);
            }
        }
    }
}
// acc[r][u] = sum_{k < n} W[k * wstride + r] * Bm[k][lane + 32 u]   (W + k * wstride 32-byte aligned)
template <int U>
__device__ __forceinline__ void ma_cols_acc(float (&acc)[kMaRows][U], const float* W, int wstride, const float* Bm, int n, int ld,
                                            int lane) {
    int col[U];      // columns beyond the tile read column 0 (discarded by the caller)
#pragma unroll
    for (int u = 0; u < U; ++u) {
        col[u] = lane + 32 * u < ld ? lane + 32 * u : 0;
#pragma unroll
        for (int r = 0; r < kMaRows; ++r) acc[r][u] = 0.f;
    }
#pragma unroll 2
    for (int k = 0; k < n; ++k) {
        const float4 w0 = *reinterpret_cast<const float4*>(W + k * wstride);
        const float4 w1 = *reinterpret_cast<const float4*>(W + k * wstride + 4);
        const float w[kMaRows] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const float bv = Bm[k * ld + col[u]];
#pragma unroll
            for (int r = 0; r < kMaRows; ++r) acc[r][u] = fmaf(w[r], bv, acc[r][u]);
        }
    }
}
// Dropout keep bits of 8 rows x ceil(L/8) groups: one Philox call per (row, group) spread over the lanes
// (round k, lane l holds pair 32 k + l = row * G + group) instead of one call per element.
struct MaKeep {
    uint32_t kp[4];
    int G;
    __device__ __forceinline__ void fill(const Dropout& dr, uint64_t row0, int L, int lane) {
        G = (L + 7) >> 3;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int pair = 32 * k + lane;
            kp[k] = 0xffu;
            if (dr.enabled() && 32 * k < kMaRows * G && pair < kMaRows * G)
                kp[k] = dr.keep8(kDropAttnProb, row0 + (uint64_t)(pair / G), (uint32_t)(pair % G));
        }
    }
    // keep bit of (row r of the step, key j); every lane of the warp must call it
    __device__ __forceinline__ bool keep(int r, int j) const {
        const int pair = r * G + (j >> 3);
        uint32_t v = 0xffu;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t x = __shfl_sync(0xffffffffu, kp[k], pair & 31);
            if ((pair >> 5) == k) v = x;
        }
        return (v >> (j & 7)) & 1u;
    }
};

template <int T, int U>
__global__ void __launch_bounds__(kMaWarps * 32) masked_attn_fwd_kernel(MaskedAttnArgs a) {
    extern __shared__ __align__(16) float sm[];
    const int L = a.L, dk = a.dk, ld = ma_ld(dk), E = a.heads * dk;
    float* Qs = sm;
    float* Ks = Qs + L * ld;
    float* Vs = Ks + L * ld;
    float* Pt = Vs + L * ld;                     // [warps][L][8]
    const int b = blockIdx.x / a.heads, h = blockIdx.x % a.heads;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* src = a.qkv + (long long)b * L * 3 * E + h * dk;
    ma_load_tile(Qs, src, 3 * E, L, dk, ld);
    ma_load_tile(Ks, src + E, 3 * E, L, dk, ld);
    ma_load_tile(Vs, src + 2 * E, 3 * E, L, dk, ld);
    ma_tiles_wait();
    __syncthreads();
    const uint8_t* mrow = a.mask ? a.mask + (long long)b * L : nullptr;
    float* pt = Pt + warp * L * kMaRows;
    bool key_real[T];
#pragma unroll
    for (int t = 0; t < T; ++t) {
        const int j = lane + 32 * t;
        key_real[t] = j < L && (!mrow || mrow[j] != 0);
    }
    for (int i0 = warp * kMaRows; i0 < L; i0 += kMaWarps * kMaRows) {
        float acc[kMaRows][T];
        ma_rows_dot<T>(acc, Qs, Ks, i0, L, ld, lane);
        const long long prow0 = ((long long)b * a.heads + h) * L + i0;
        MaKeep mk;
        mk.fill(a.drop, (uint64_t)prow0, L, lane);
#pragma unroll
        for (int r = 0; r < kMaRows; ++r) {
            const int i = i0 + r;
            const bool row_ok = i < L;
            const bool qi_real = !mrow || mrow[min(i, L - 1)] != 0;
            float mx = -INFINITY;
#pragma unroll
            for (int t = 0; t < T; ++t) {
                    float v = acc[r][t] * a.scale;
                    if (!(qi_real && key_real[t])) v = -1e9f;      // masked_fill (nrms.py:38-41)
                    acc[r][t] = v;
                    if (lane + 32 * t < L) mx = fmaxf(mx, v);
                }
            mx = warp_max(mx);
            float sum = 0.f;
#pragma unroll
            for (int t = 0; t < T; ++t) {
                    acc[r][t] = lane + 32 * t < L ? expf(acc[r][t] - mx) : 0.f;
                    sum += acc[r][t];
                }
            sum = warp_sum(sum);
            const float inv = 1.f / sum;
#pragma unroll
            for (int t = 0; t < T; ++t) {
                    const int j = lane + 32 * t;
                    float p = acc[r][t] * inv;
                    const bool kept = mk.keep(r, min(j, L - 1));
                    if (j < L) {
                        if (row_ok) a.probs[(prow0 + r) * L + j] = p;
                        if (a.drop.enabled()) p = kept ? p * a.drop.scale : 0.f;
                        pt[j * kMaRows + r] = p;
                    }
                }
        }
        __syncwarp();
        float o[kMaRows][U];
        ma_cols_acc<U>(o, pt, kMaRows, Vs, L, ld, lane);
#pragma unroll
        for (int r = 0; r < kMaRows; ++r)
            if (i0 + r < L) {
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int d = lane + 32 * u;
                    if (d < dk) a.ctx[((long long)b * L + i0 + r) * E + h * dk + d] = o[r][u];
                }
            }
        __syncwarp();
    }
}

// backward of the above: dV = Pd^T dO, dPd = dO V^T, dP = dPd * keep/(1-p), dS = P (dP - rowsum(P dP)) / sqrt(dk)
// with dS = 0 wherever the score was overwritten by the mask (masked_fill passes no gradient), dQ = dS K, dK = dS^T Q.
template <int T, int U>
__global__ void __launch_bounds__(kMaWarps * 32) masked_attn_bwd_kernel(MaskedAttnArgs a) {
    extern __shared__ __align__(16) float sm[];
    const int L = a.L, dk = a.dk, ld = ma_ld(dk), E = a.heads * dk, lp = ma_lp(L);
    float* Qs = sm;
    float* Ks = Qs + L * ld;
    float* Vs = Ks + L * ld;
    float* Os = Vs + L * ld;                     // dO
    float* dS = Os + L * ld;                     // [L][lp]
    float* Pd = dS + L * lp;                     // [L][lp] dropped probabilities
    float* St = Pd + L * lp;                     // [warps][L][8]: a warp's dS rows, transposed
    const int b = blockIdx.x / a.heads, h = blockIdx.x % a.heads;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* src = a.qkv + (long long)b * L * 3 * E + h * dk;
    ma_load_tile(Qs, src, 3 * E, L, dk, ld);
    ma_load_tile(Ks, src + E, 3 * E, L, dk, ld);
    ma_load_tile(Vs, src + 2 * E, 3 * E, L, dk, ld);
    ma_load_tile(Os, a.d_ctx + (long long)b * L * E + h * dk, E, L, dk, ld);
    for (int i = threadIdx.x; i < L * (lp - L); i += blockDim.x) {      // padding columns: read (and discarded) by dK / dV
        const int r = i / (lp - L), c = L + i % (lp - L);
        dS[r * lp + c] = 0.f;
        Pd[r * lp + c] = 0.f;
    }
    ma_tiles_wait();
    __syncthreads();
    const uint8_t* mrow = a.mask ? a.mask + (long long)b * L : nullptr;
    float* st = St + warp * L * kMaRows;
    float* dst = a.d_qkv + (long long)b * L * 3 * E + h * dk;
    bool key_real[T];
#pragma unroll
    for (int t = 0; t < T; ++t) {
        const int j = lane + 32 * t;
        key_real[t] = j < L && (!mrow || mrow[j] != 0);
    }
    for (int i0 = warp * kMaRows; i0 < L; i0 += kMaWarps * kMaRows) {
        float acc[kMaRows][T];
        ma_rows_dot<T>(acc, Os, Vs, i0, L, ld, lane);                   // dPd
        const long long prow0 = ((long long)b * a.heads + h) * L + i0;
        MaKeep mk;
        mk.fill(a.drop, (uint64_t)prow0, L, lane);
#pragma unroll
        for (int r = 0; r < kMaRows; ++r) {
            const int i = min(i0 + r, L - 1);
            const bool row_ok = i0 + r < L;
            const bool qi_real = !mrow || mrow[i] != 0;
            float p[T];
            float delta = 0.f;
#pragma unroll
            for (int t = 0; t < T; ++t) {
                    const int j = lane + 32 * t;
                    const bool kept = mk.keep(r, min(j, L - 1));
                    p[t] = 0.f;
                    float dp = 0.f;
                    if (j < L) {
                        p[t] = a.probs[(prow0 - i0 + i) * L + j];
                        const float mult = a.drop.enabled() ? (kept ? a.drop.scale : 0.f) : 1.f;
                        if (row_ok) Pd[i * lp + j] = p[t] * mult;
                        dp = acc[r][t] * mult;
                    }
                    acc[r][t] = dp;
                    delta = fmaf(p[t], dp, delta);
                }
            delta = warp_sum(delta);
#pragma unroll
            for (int t = 0; t < T; ++t) {
                    const int j = lane + 32 * t;
                    if (j < L) {
                        const float v = (qi_real && key_real[t]) ? p[t] * (acc[r][t] - delta) * a.scale : 0.f;
                        if (row_ok) dS[i * lp + j] = v;
                        st[j * kMaRows + r] = v;
                    }
                }
        }
        __syncwarp();
        float o[kMaRows][U];
        ma_cols_acc<U>(o, st, kMaRows, Ks, L, ld, lane);                // dQ rows [i0, i0 + 8)
#pragma unroll
        for (int r = 0; r < kMaRows; ++r)
            if (i0 + r < L) {
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int d = lane + 32 * u;
                    if (d < dk) dst[(long long)(i0 + r) * 3 * E + d] = o[r][u];
                }
            }
        __syncwarp();
    }
    __syncthreads();
    for (int j0 = warp * kMaRows; j0 < L; j0 += kMaWarps * kMaRows) {
        float o[kMaRows][U];
        ma_cols_acc<U>(o, dS + j0, lp, Qs, L, ld, lane);                // dK rows [j0, j0 + 8): sum over query rows
#pragma unroll
        for (int r = 0; r < kMaRows; ++r)
            if (j0 + r < L) {
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int d = lane + 32 * u;
                    if (d < dk) dst[(long long)(j0 + r) * 3 * E + E + d] = o[r][u];
                }
            }
        ma_cols_acc<U>(o, Pd + j0, lp, Os, L, ld, lane);                // dV
#pragma unroll
        for (int r = 0; r < kMaRows; ++r)
            if (j0 + r < L) {
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int d = lane + 32 * u;
                    if (d < dk) dst[(long long)(j0 + r) * 3 * E + 2 * E + d] = o[r][u];
                }
            }
    }
}

// keys per lane T = ceil(L / 32) and head columns per lane U = ceil(dk / 32) are compile-time (fully unrolled
// register tiles); 3 rounds up to 4
template <bool BWD>
cudaError_t launch_masked_attn(const MaskedAttnArgs& a, size_t smem, cudaStream_t s) {
    const int T = ceil_div(a.L, 32), U = ceil_div(a.dk, 32);
#define NRMS_MA_CASE(TT, UU)                                                                                       \
    if (T <= TT && U <= UU) {                                                                                      \
        auto k = BWD ? masked_attn_bwd_kernel<TT, UU> : masked_attn_fwd_kernel<TT, UU>;                      \
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);            \
        if (e != cudaSuccess) return e;                                                                            \
        NRMS_LAUNCH(BWD ? "masked_attn_bwd" : "masked_attn_fwd", s, k<<<a.B * a.heads, kMaWarps * 32, smem, s>>>(a)); \
        return cudaGetLastError();                                                                                 \
    }
    NRMS_MA_CASE(1, 1) NRMS_MA_CASE(1, 2) NRMS_MA_CASE(1, 4)
    NRMS_MA_CASE(2, 1) NRMS_MA_CASE(2, 2) NRMS_MA_CASE(2, 4)
    NRMS_MA_CASE(4, 1) NRMS_MA_CASE(4, 2) NRMS_MA_CASE(4, 4)
#undef NRMS_MA_CASE
    return cudaErrorInvalidValue;
}

// ---- masked additive attention (nrms.py:98-117) after its Linear ------------------------------------------
// t [B, L, Q] = x W^T + b (pre-tanh), qv [Q], x [B, L, E], mask [B, L] or NULL -> alpha [B, L] (softmax
// weights), out [B, E].  One CTA per sequence.
struct MaskedPoolArgs {
    const float* t;
    const float* qv;
    const float* x;
    const uint8_t* mask;
    float* alpha;
    float* out;
    const float* d_out;   // backward: [B, E]
    float* d_t;           // backward: [B, L, Q]
    float* d_x;           // backward: [B, L, E]  (the pooling's own share: alpha_i * d_out)
    float* d_qv_part;     // backward: [B, Q] per-sequence partials of the query-vector gradient
    int B, L, Q, E;
};
inline size_t masked_pool_smem(int L) { return sizeof(float) * (size_t)(2 * L + 32); }

__global__ void __launch_bounds__(256) masked_pool_fwd_kernel(MaskedPoolArgs a) {
    extern __shared__ float sm[];
    float* sc = sm;                 // [L] scores -> weights
    const int b = blockIdx.x, L = a.L, Q = a.Q, E = a.E;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* t = a.t + (long long)b * L * Q;
    for (int i = warp; i < L; i += 8) {
        float s = 0.f;
        for (int q = lane; q < Q; q += 32) s = fmaf(tanhf(t[(long long)i * Q + q]), a.qv[q], s);
        s = warp_sum(s);
        if (lane == 0) sc[i] = (a.mask && a.mask[(long long)b * L + i] == 0) ? -1e9f : s;
    }
    __syncthreads();
    if (warp == 0) {
        float mx = -INFINITY;
        for (int i = lane; i < L; i += 32) mx = fmaxf(mx, sc[i]);
        mx = warp_max(mx);
        float sum = 0.f;
        for (int i = lane; i < L; i += 32) {
            const float e = expf(sc[i] - mx);
            sc[i] = e;
            sum += e;
        }
        sum = warp_sum(sum);
        const float inv = 1.f / sum;
        for (int i = lane; i < L; i += 32) {
            sc[i] *= inv;
            a.alpha[(long long)b * L + i] = sc[i];
        }
    }
    __syncthreads();
    const float* x = a.x + (long long)b * L * E;
    for (int e = threadIdx.x; e < E; e += blockDim.x) {
        float acc = 0.f;
        for (int i = 0; i < L; ++i) acc = fmaf(sc[i], x[(long long)i * E + e], acc);
        a.out[(long long)b * E + e] = acc;
    }
}

__global__ void __launch_bounds__(256) masked_pool_bwd_kernel(MaskedPoolArgs a) {
    extern __shared__ float sm[];
    float* dw = sm;                 // [L] d(weight) -> d(score)
    float* al = sm + a.L;           // [L]
    float* red = al + a.L;          // [32]
    const int b = blockIdx.x, L = a.L, Q = a.Q, E = a.E;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* x = a.x + (long long)b * L * E;
    const float* go = a.d_out + (long long)b * E;
    for (int i = warp; i < L; i += 8) {
        float s = 0.f;
        for (int e = lane; e < E; e += 32) s = fmaf(go[e], x[(long long)i * E + e], s);
        s = warp_sum(s);
        if (lane == 0) {
            dw[i] = s;
            al[i] = a.alpha[(long long)b * L + i];
        }
    }
    __syncthreads();
    if (warp == 0) {
        float dot = 0.f;
        for (int i = lane; i < L; i += 32) dot = fmaf(al[i], dw[i], dot);
        dot = warp_sum(dot);
        if (lane == 0) red[0] = dot;
    }
    __syncthreads();
    const float dot = red[0];
    __syncthreads();
    for (int i = threadIdx.x; i < L; i += blockDim.x) {
        const bool live = !a.mask || a.mask[(long long)b * L + i] != 0;     // masked_fill passes no gradient
        dw[i] = live ? al[i] * (dw[i] - dot) : 0.f;
    }
    __syncthreads();
    // d_x (pooling share), d_t
    float* dx = a.d_x + (long long)b * L * E;
    for (long long k = threadIdx.x; k < (long long)L * E; k += blockDim.x) {
        const int i = (int)(k / E), e = (int)(k - (long long)i * E);
        dx[k] = al[i] * go[e];
    }
    const float* t = a.t + (long long)b * L * Q;
    float* dt = a.d_t + (long long)b * L * Q;
    for (int q = threadIdx.x; q < Q; q += blockDim.x) {
        const float qv = a.qv[q];
        float dq = 0.f;
        for (int i = 0; i < L; ++i) {
            const float th = tanhf(t[(long long)i * Q + q]);
            dt[(long long)i * Q + q] = dw[i] * qv * (1.f - th * th);
            dq = fmaf(dw[i], th, dq);
        }
        a.d_qv_part[(long long)b * Q + q] = dq;
    }
}

}  // namespace mu
}  // namespace nrms
