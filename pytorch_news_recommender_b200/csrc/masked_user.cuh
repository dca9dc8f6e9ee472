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
// ctx [B, L, E].  One CTA per (sequence, head); K, V and Q of the item in shared memory ([L][dk + 1] fp32),
// a warp per query row: lane j owns keys j, j + 32, ...; the row's probabilities go through a per-warp
// shared row for the P.V product (lane d owns columns d, d + 32, ...).
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
constexpr int kMaMaxKeysPerLane = 4;      // L <= 128

inline size_t masked_attn_fwd_smem(int L, int dk) {
    return sizeof(float) * ((size_t)3 * L * (dk + 1) + (size_t)kMaWarps * L);
}
inline size_t masked_attn_bwd_smem(int L, int dk) {
    return sizeof(float) * ((size_t)4 * L * (dk + 1) + (size_t)2 * L * (L + 1));
}

__global__ void __launch_bounds__(kMaWarps * 32) masked_attn_fwd_kernel(MaskedAttnArgs a) {
    extern __shared__ float sm[];
    const int L = a.L, dk = a.dk, ld = dk + 1, E = a.heads * dk;
    float* Qs = sm;
    float* Ks = Qs + L * ld;
    float* Vs = Ks + L * ld;
    float* Ps = Vs + L * ld;                     // [warps][L]
    const int b = blockIdx.x / a.heads, h = blockIdx.x % a.heads;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* src = a.qkv + (long long)b * L * 3 * E + h * dk;
    for (int i = threadIdx.x; i < L * dk; i += blockDim.x) {
        const int r = i / dk, d = i - r * dk;
        const float* p = src + (long long)r * 3 * E + d;
        Qs[r * ld + d] = p[0];
        Ks[r * ld + d] = p[E];
        Vs[r * ld + d] = p[2 * E];
    }
    __syncthreads();
    const uint8_t* mrow = a.mask ? a.mask + (long long)b * L : nullptr;
    float* prow = Ps + warp * L;
    for (int i = warp; i < L; i += kMaWarps) {
        const bool qi_real = !mrow || mrow[i] != 0;
        float s[kMaMaxKeysPerLane];
        float mx = -INFINITY;
#pragma unroll
        for (int t = 0; t < kMaMaxKeysPerLane; ++t) {
            const int j = lane + 32 * t;
            float acc = 0.f;
            if (j < L) {
                for (int d = 0; d < dk; ++d) acc = fmaf(Qs[i * ld + d], Ks[j * ld + d], acc);
                acc *= a.scale;
                if (!(qi_real && (!mrow || mrow[j] != 0))) acc = -1e9f;      // masked_fill (nrms.py:38-41)
                mx = fmaxf(mx, acc);
            }
            s[t] = acc;
        }
        mx = warp_max(mx);
        float sum = 0.f;
#pragma unroll
        for (int t = 0; t < kMaMaxKeysPerLane; ++t) {
            const int j = lane + 32 * t;
            s[t] = j < L ? expf(s[t] - mx) : 0.f;
            sum += s[t];
        }
        sum = warp_sum(sum);
        const float inv = 1.f / sum;
        const long long prow_g = (((long long)b * a.heads + h) * L + i);
#pragma unroll
        for (int t = 0; t < kMaMaxKeysPerLane; ++t) {
            const int j = lane + 32 * t;
            if (j < L) {
                float p = s[t] * inv;
                a.probs[prow_g * L + j] = p;
                if (a.drop.enabled()) {
                    const uint32_t keep = a.drop.keep8(kDropAttnProb, (uint64_t)prow_g, (uint32_t)(j >> 3));
                    p = ((keep >> (j & 7)) & 1u) ? p * a.drop.scale : 0.f;
                }
                prow[j] = p;
            }
        }
        __syncwarp();
        for (int d = lane; d < dk; d += 32) {
            float acc = 0.f;
            for (int j = 0; j < L; ++j) acc = fmaf(prow[j], Vs[j * ld + d], acc);
            a.ctx[((long long)b * L + i) * E + h * dk + d] = acc;
        }
        __syncwarp();
    }
}

// backward of the above: dV = Pd^T dO, dPd = dO V^T, dP = dPd * keep/(1-p), dS = P (dP - rowsum(P dP)) / sqrt(dk)
// with dS = 0 wherever the score was overwritten by the mask (masked_fill passes no gradient), dQ = dS K, dK = dS^T Q.
__global__ void __launch_bounds__(kMaWarps * 32) masked_attn_bwd_kernel(MaskedAttnArgs a) {
    extern __shared__ float sm[];
    const int L = a.L, dk = a.dk, ld = dk + 1, E = a.heads * dk, lp = L + 1;
    float* Qs = sm;
    float* Ks = Qs + L * ld;
    float* Vs = Ks + L * ld;
    float* Os = Vs + L * ld;                     // dO
    float* dS = Os + L * ld;                     // [L][L + 1]
    float* Pd = dS + L * lp;                     // [L][L + 1] dropped probabilities
    const int b = blockIdx.x / a.heads, h = blockIdx.x % a.heads;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* src = a.qkv + (long long)b * L * 3 * E + h * dk;
    const float* dsrc = a.d_ctx + (long long)b * L * E + h * dk;
    for (int i = threadIdx.x; i < L * dk; i += blockDim.x) {
        const int r = i / dk, d = i - r * dk;
        const float* p = src + (long long)r * 3 * E + d;
        Qs[r * ld + d] = p[0];
        Ks[r * ld + d] = p[E];
        Vs[r * ld + d] = p[2 * E];
        Os[r * ld + d] = dsrc[(long long)r * E + d];
    }
    __syncthreads();
    const uint8_t* mrow = a.mask ? a.mask + (long long)b * L : nullptr;
    for (int i = warp; i < L; i += kMaWarps) {
        const bool qi_real = !mrow || mrow[i] != 0;
        const long long prow_g = (((long long)b * a.heads + h) * L + i);
        float p[kMaMaxKeysPerLane], dp[kMaMaxKeysPerLane];
        float delta = 0.f;
#pragma unroll
        for (int t = 0; t < kMaMaxKeysPerLane; ++t) {
            const int j = lane + 32 * t;
            p[t] = 0.f;
            dp[t] = 0.f;
            if (j < L) {
                float acc = 0.f;
                for (int d = 0; d < dk; ++d) acc = fmaf(Os[i * ld + d], Vs[j * ld + d], acc);
                p[t] = a.probs[prow_g * L + j];
                float mult = 1.f;
                if (a.drop.enabled()) {
                    const uint32_t keep = a.drop.keep8(kDropAttnProb, (uint64_t)prow_g, (uint32_t)(j >> 3));
                    mult = ((keep >> (j & 7)) & 1u) ? a.drop.scale : 0.f;
                }
                Pd[i * lp + j] = p[t] * mult;
                dp[t] = acc * mult;
                delta = fmaf(p[t], dp[t], delta);
            }
        }
        delta = warp_sum(delta);
#pragma unroll
        for (int t = 0; t < kMaMaxKeysPerLane; ++t) {
            const int j = lane + 32 * t;
            if (j < L) {
                const bool live = qi_real && (!mrow || mrow[j] != 0);
                dS[i * lp + j] = live ? p[t] * (dp[t] - delta) * a.scale : 0.f;
            }
        }
    }
    __syncthreads();
    float* dst = a.d_qkv + (long long)b * L * 3 * E + h * dk;
    for (int r = warp; r < L; r += kMaWarps) {
        for (int d = lane; d < dk; d += 32) {
            float dq = 0.f, dkk = 0.f, dv = 0.f;
            for (int j = 0; j < L; ++j) {
                dq = fmaf(dS[r * lp + j], Ks[j * ld + d], dq);          // row r as a query
                dkk = fmaf(dS[j * lp + r], Qs[j * ld + d], dkk);        // row r as a key
                dv = fmaf(Pd[j * lp + r], Os[j * ld + d], dv);
            }
            float* o = dst + (long long)r * 3 * E + d;
            o[0] = dq;
            o[E] = dkk;
            o[2 * E] = dv;
        }
    }
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
