// attention.cuh — per-head scaled-dot-product self-attention over short sequences
// (title: 20-48 words, history: 50-200 news), forward and backward.
// Reference: ScaledDotProductAttention.forward nrms_v0.py:13-23 called from
// MultiHeadSelfAttention.forward nrms_v0.py:46-76 — softmax(Q K^T / sqrt(d_k)) V, NO mask
// (length=None always, nrms_v0.py:170,196), NO output projection.
//
// Layout: qkv [M, 3D] rows = tokens, [Q | K | V], head h owns columns h*dk .. h*dk+dk-1 of
// each third (the view/transpose of nrms_v0.py:53-58).  One CTA handles one sequence and a
// group of `hpb` heads; one warp handles (head, 32-row chunk); one LANE owns one query row
// and streams over the keys held in shared memory (broadcast float4 reads) with an online
// softmax, so scores never leave registers.  d_k <= 32 (30 in the reference config).
#pragma once
#include "common.cuh"

namespace nrms {

constexpr int kDkPad = 32;   // per-head dim padded to 32 (pad lanes are zero)
constexpr int kRowStride = 36;  // smem row stride in floats: float4-aligned and conflict-free
                                // both for broadcast reads and lane-per-row float4 reads

struct AttnArgs {
    const float* qkv;   // [M, 3D]
    float* ctx;         // fwd out / bwd in (post-dropout context) [M, D]
    float* lse;         // [M, n_heads] log-sum-exp of the scaled scores
    const float* d_ctx; // bwd: grad wrt post-dropout context [M, D]
    float* d_qkv;       // bwd out [M, 3D]
    float* d_bias_part; // bwd out [n_seq, 3D]: per-sequence column sums of d_qkv
    int L, D, n_heads, dk, hpb;
    float scale;        // 1/sqrt(dk)
    Dropout drop;       // context dropout (stream kDropContext), disabled in eval / user encoder
};

__host__ __device__ inline size_t attn_fwd_smem_bytes(int L, int hpb) {
    return (size_t)3 * hpb * L * kRowStride * sizeof(float);
}
__host__ __device__ inline size_t attn_bwd_smem_bytes(int L, int hpb) {
    return (size_t)4 * hpb * L * kRowStride * sizeof(float);
}

// cooperative load of one of the three thirds (or d_ctx) for the CTA's heads into smem
// dst[(hh*L + l)*kRowStride + d], zero padded for d in [dk, 32).
__device__ __forceinline__ void load_heads(float* dst, const float* src, long long row0,
                                           int ld, int col0, int L, int dk, int hpb,
                                           float mul) {
    const int ncol = hpb * dk;
    for (int i = threadIdx.x; i < L * ncol; i += blockDim.x) {
        const int l = i / ncol, c = i - l * ncol;
        const int hh = c / dk, d = c - hh * dk;
        dst[(hh * L + l) * kRowStride + d] = __ldg(src + (row0 + l) * ld + col0 + c) * mul;
    }
    const int npad = kDkPad - dk;
    if (npad > 0) {
        for (int i = threadIdx.x; i < hpb * L * npad; i += blockDim.x) {
            const int r = i / npad, d = dk + (i - r * npad);
            dst[r * kRowStride + d] = 0.f;
        }
    }
}

__global__ void __launch_bounds__(512) attn_fwd_kernel(const AttnArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int L = a.L, D = a.D, dk = a.dk;
    const int seq = blockIdx.x;
    const int h0 = blockIdx.y * a.hpb;
    const int hpb = min(a.hpb, a.n_heads - h0);
    float* Qs = smem;
    float* Ks = Qs + (size_t)a.hpb * L * kRowStride;
    float* Vs = Ks + (size_t)a.hpb * L * kRowStride;
    const long long row0 = (long long)seq * L;
    const int ld = 3 * D;

    load_heads(Qs, a.qkv, row0, ld, h0 * dk, L, dk, hpb, a.scale);
    load_heads(Ks, a.qkv, row0, ld, D + h0 * dk, L, dk, hpb, 1.f);
    load_heads(Vs, a.qkv, row0, ld, 2 * D + h0 * dk, L, dk, hpb, 1.f);
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rc = ceil_div(L, 32);
    const int nwarps = blockDim.x >> 5;
    for (int task = warp; task < hpb * rc; task += nwarps) {
        const int hh = task / rc;
        const int i = (task - hh * rc) * 32 + lane;
        const bool active = i < L;
        const int ii = active ? i : 0;
        float q[kDkPad], acc[kDkPad];
        const float* qrow = Qs + (size_t)(hh * L + ii) * kRowStride;
#pragma unroll
        for (int c = 0; c < kDkPad / 4; ++c) {
            const float4 v = *reinterpret_cast<const float4*>(qrow + 4 * c);
            q[4 * c] = v.x; q[4 * c + 1] = v.y; q[4 * c + 2] = v.z; q[4 * c + 3] = v.w;
        }
#pragma unroll
        for (int d = 0; d < kDkPad; ++d) acc[d] = 0.f;
        float mx = -INFINITY, den = 0.f;
        const float* kbase = Ks + (size_t)hh * L * kRowStride;
        const float* vbase = Vs + (size_t)hh * L * kRowStride;
        for (int j = 0; j < L; ++j) {
            const float* kr = kbase + j * kRowStride;
            float s = 0.f;
#pragma unroll
            for (int c = 0; c < kDkPad / 4; ++c) {
                const float4 kv = *reinterpret_cast<const float4*>(kr + 4 * c);
                s = fmaf(q[4 * c], kv.x, s);
                s = fmaf(q[4 * c + 1], kv.y, s);
                s = fmaf(q[4 * c + 2], kv.z, s);
                s = fmaf(q[4 * c + 3], kv.w, s);
            }
            const float mnew = fmaxf(mx, s);
            const float corr = __expf(mx - mnew);  // exp(-inf)=0 on the first key
            const float p = __expf(s - mnew);
            den = den * corr + p;
            const float* vr = vbase + j * kRowStride;
#pragma unroll
            for (int c = 0; c < kDkPad / 4; ++c) {
                const float4 vv = *reinterpret_cast<const float4*>(vr + 4 * c);
                acc[4 * c] = fmaf(p, vv.x, acc[4 * c] * corr);
                acc[4 * c + 1] = fmaf(p, vv.y, acc[4 * c + 1] * corr);
                acc[4 * c + 2] = fmaf(p, vv.z, acc[4 * c + 2] * corr);
                acc[4 * c + 3] = fmaf(p, vv.w, acc[4 * c + 3] * corr);
            }
            mx = mnew;
        }
        if (active) {
            const float inv = 1.f / den;
            float* orow = Qs + (size_t)(hh * L + i) * kRowStride;  // own row: safe to overwrite
#pragma unroll
            for (int c = 0; c < kDkPad / 4; ++c)
                *reinterpret_cast<float4*>(orow + 4 * c) = make_float4(
                    acc[4 * c] * inv, acc[4 * c + 1] * inv, acc[4 * c + 2] * inv,
                    acc[4 * c + 3] * inv);
            a.lse[(row0 + i) * a.n_heads + h0 + hh] = mx + __logf(den);
        }
    }
    __syncthreads();
    // coalesced write-out (+ context dropout, nrms_v0.py:171-173)
    const int ncol = hpb * dk;
    for (int i = threadIdx.x; i < L * ncol; i += blockDim.x) {
        const int l = i / ncol, c = i - l * ncol;
        const int hh = c / dk, d = c - hh * dk;
        float v = Qs[(size_t)(hh * L + l) * kRowStride + d];
        const long long e = (row0 + l) * D + h0 * dk + c;
        if (a.drop.enabled()) v *= a.drop.mult(kDropContext, (uint64_t)e);
        a.ctx[e] = v;
    }
}

// Backward.  With P = softmax(S), S = scale * Q K^T, O = P V, dO given:
//   dV = P^T dO ; dP = dO V^T ; dS = P * (dP - rowsum(dO*O)) ; dQ = scale dS K ; dK = scale dS^T Q
// Pass A: lane = query row i (accumulates dQ_i); pass B: lane = key row j (accumulates
// dK_j, dV_j).  P is recomputed from the saved log-sum-exp.
__global__ void __launch_bounds__(256) attn_bwd_kernel(const AttnArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int L = a.L, D = a.D, dk = a.dk;
    const int seq = blockIdx.x;
    const int h0 = blockIdx.y * a.hpb;
    const int hpb = min(a.hpb, a.n_heads - h0);
    const size_t per = (size_t)a.hpb * L * kRowStride;
    float* Qs = smem;  // holds scale*Q
    float* Ks = Qs + per;
    float* Vs = Ks + per;
    float* Gs = Vs + per;  // dO (grad wrt pre-dropout context)
    const long long row0 = (long long)seq * L;
    const int ld = 3 * D;

    load_heads(Qs, a.qkv, row0, ld, h0 * dk, L, dk, hpb, a.scale);
    load_heads(Ks, a.qkv, row0, ld, D + h0 * dk, L, dk, hpb, 1.f);
    load_heads(Vs, a.qkv, row0, ld, 2 * D + h0 * dk, L, dk, hpb, 1.f);
    {
        const int ncol = hpb * dk;
        for (int i = threadIdx.x; i < L * ncol; i += blockDim.x) {
            const int l = i / ncol, c = i - l * ncol;
            const int hh = c / dk, d = c - hh * dk;
            const long long e = (row0 + l) * D + h0 * dk + c;
            float v = __ldg(a.d_ctx + e);
            if (a.drop.enabled()) v *= a.drop.mult(kDropContext, (uint64_t)e);
            Gs[(size_t)(hh * L + l) * kRowStride + d] = v;
        }
        // column 31 of every Gs row carries delta_i = rowsum(dO*O) (dk <= 30) or it is
        // kept in a register when dk > 30 -> we always recompute it per lane below and
        // broadcast through column `kDkPad` .. kRowStride-1 (4 spare floats per row).
        const int npad = kDkPad - dk;
        if (npad > 0)
            for (int i = threadIdx.x; i < hpb * L * npad; i += blockDim.x) {
                const int r = i / npad, d = dk + (i - r * npad);
                Gs[(size_t)r * kRowStride + d] = 0.f;
            }
    }
    __syncthreads();
    // delta_i and lse_i go to the spare columns 32,33 of Gs rows
    for (int i = threadIdx.x; i < hpb * L; i += blockDim.x) {
        const int hh = i / L, l = i - hh * L;
        const float* g = Gs + (size_t)i * kRowStride;
        const float* o = a.ctx + (row0 + l) * D + (h0 + hh) * dk;  // post-dropout context
        const float* go = a.d_ctx + (row0 + l) * D + (h0 + hh) * dk;
        float dl = 0.f;
        // sum(dO_raw * O_raw) == sum(dO_out * O_out) because both carry the same mask factor
        for (int d = 0; d < dk; ++d) dl = fmaf(__ldg(go + d), __ldg(o + d), dl);
        (void)g;
        Gs[(size_t)i * kRowStride + kDkPad] = dl;
        Gs[(size_t)i * kRowStride + kDkPad + 1] = a.lse[(row0 + l) * a.n_heads + h0 + hh];
    }
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rc = ceil_div(L, 32);
    const int nwarps = blockDim.x >> 5;
    // results are staged in registers until every warp has finished reading smem
    // (tasks per warp is small: hpb*rc / nwarps, launch config guarantees exactly 1)
    const int task = warp;
    const bool has_task = task < hpb * rc;
    const int hh = has_task ? task / rc : 0;
    const int i = has_task ? (task - hh * rc) * 32 + lane : 0;
    const bool active = has_task && i < L;
    const int ii = active ? i : 0;
    const float* qb = Qs + (size_t)hh * L * kRowStride;
    const float* kb = Ks + (size_t)hh * L * kRowStride;
    const float* vb = Vs + (size_t)hh * L * kRowStride;
    const float* gb = Gs + (size_t)hh * L * kRowStride;

    float dq[kDkPad];
    {
        // ---- pass A: lane = query row ---------------------------------------------------
        float q[kDkPad], go[kDkPad];
#pragma unroll
        for (int c = 0; c < kDkPad / 4; ++c) {
            const float4 v = *reinterpret_cast<const float4*>(qb + ii * kRowStride + 4 * c);
            q[4 * c] = v.x; q[4 * c + 1] = v.y; q[4 * c + 2] = v.z; q[4 * c + 3] = v.w;
            const float4 w = *reinterpret_cast<const float4*>(gb + ii * kRowStride + 4 * c);
            go[4 * c] = w.x; go[4 * c + 1] = w.y; go[4 * c + 2] = w.z; go[4 * c + 3] = w.w;
        }
        const float delta = gb[ii * kRowStride + kDkPad];
        const float lse = gb[ii * kRowStride + kDkPad + 1];
#pragma unroll
        for (int d = 0; d < kDkPad; ++d) dq[d] = 0.f;
        if (has_task) {
            for (int j = 0; j < L; ++j) {
                const float* kr = kb + j * kRowStride;
                const float* vr = vb + j * kRowStride;
                float s = 0.f, dp = 0.f;
                float kreg[kDkPad];
#pragma unroll
                for (int c = 0; c < kDkPad / 4; ++c) {
                    const float4 kv = *reinterpret_cast<const float4*>(kr + 4 * c);
                    const float4 vv = *reinterpret_cast<const float4*>(vr + 4 * c);
                    kreg[4 * c] = kv.x; kreg[4 * c + 1] = kv.y; kreg[4 * c + 2] = kv.z;
                    kreg[4 * c + 3] = kv.w;
                    s = fmaf(q[4 * c], kv.x, s); s = fmaf(q[4 * c + 1], kv.y, s);
                    s = fmaf(q[4 * c + 2], kv.z, s); s = fmaf(q[4 * c + 3], kv.w, s);
                    dp = fmaf(go[4 * c], vv.x, dp); dp = fmaf(go[4 * c + 1], vv.y, dp);
                    dp = fmaf(go[4 * c + 2], vv.z, dp); dp = fmaf(go[4 * c + 3], vv.w, dp);
                }
                const float p = __expf(s - lse);
                const float ds = p * (dp - delta) * a.scale;
#pragma unroll
                for (int d = 0; d < kDkPad; ++d) dq[d] = fmaf(ds, kreg[d], dq[d]);
            }
        }
    }
    float dkk[kDkPad], dvv[kDkPad];
    {
        // ---- pass B: lane = key row -------------------------------------------------------
        float k[kDkPad], v[kDkPad];
#pragma unroll
        for (int c = 0; c < kDkPad / 4; ++c) {
            const float4 x = *reinterpret_cast<const float4*>(kb + ii * kRowStride + 4 * c);
            k[4 * c] = x.x; k[4 * c + 1] = x.y; k[4 * c + 2] = x.z; k[4 * c + 3] = x.w;
            const float4 y = *reinterpret_cast<const float4*>(vb + ii * kRowStride + 4 * c);
            v[4 * c] = y.x; v[4 * c + 1] = y.y; v[4 * c + 2] = y.z; v[4 * c + 3] = y.w;
        }
#pragma unroll
        for (int d = 0; d < kDkPad; ++d) { dkk[d] = 0.f; dvv[d] = 0.f; }
        if (has_task) {
            for (int r = 0; r < L; ++r) {
                const float* qr = qb + r * kRowStride;  // scale*Q_r
                const float* gr = gb + r * kRowStride;
                float s = 0.f, dp = 0.f;
                float qreg[kDkPad], greg[kDkPad];
#pragma unroll
                for (int c = 0; c < kDkPad / 4; ++c) {
                    const float4 qv = *reinterpret_cast<const float4*>(qr + 4 * c);
                    const float4 gv = *reinterpret_cast<const float4*>(gr + 4 * c);
                    qreg[4 * c] = qv.x; qreg[4 * c + 1] = qv.y; qreg[4 * c + 2] = qv.z;
                    qreg[4 * c + 3] = qv.w;
                    greg[4 * c] = gv.x; greg[4 * c + 1] = gv.y; greg[4 * c + 2] = gv.z;
                    greg[4 * c + 3] = gv.w;
                    s = fmaf(qv.x, k[4 * c], s); s = fmaf(qv.y, k[4 * c + 1], s);
                    s = fmaf(qv.z, k[4 * c + 2], s); s = fmaf(qv.w, k[4 * c + 3], s);
                    dp = fmaf(gv.x, v[4 * c], dp); dp = fmaf(gv.y, v[4 * c + 1], dp);
                    dp = fmaf(gv.z, v[4 * c + 2], dp); dp = fmaf(gv.w, v[4 * c + 3], dp);
                }
                const float delta = gr[kDkPad], lse = gr[kDkPad + 1];
                const float p = __expf(s - lse);
                const float ds = p * (dp - delta);  // qreg already carries `scale`
#pragma unroll
                for (int d = 0; d < kDkPad; ++d) {
                    dvv[d] = fmaf(p, greg[d], dvv[d]);
                    dkk[d] = fmaf(ds, qreg[d], dkk[d]);
                }
            }
        }
    }
    __syncthreads();  // everyone is done reading Q/K/V/dO
    if (active) {
        float* r0 = Qs + (size_t)(hh * L + i) * kRowStride;
        float* r1 = Ks + (size_t)(hh * L + i) * kRowStride;
        float* r2 = Vs + (size_t)(hh * L + i) * kRowStride;
#pragma unroll
        for (int c = 0; c < kDkPad / 4; ++c) {
            *reinterpret_cast<float4*>(r0 + 4 * c) =
                make_float4(dq[4 * c], dq[4 * c + 1], dq[4 * c + 2], dq[4 * c + 3]);
            *reinterpret_cast<float4*>(r1 + 4 * c) =
                make_float4(dkk[4 * c], dkk[4 * c + 1], dkk[4 * c + 2], dkk[4 * c + 3]);
            *reinterpret_cast<float4*>(r2 + 4 * c) =
                make_float4(dvv[4 * c], dvv[4 * c + 1], dvv[4 * c + 2], dvv[4 * c + 3]);
        }
    }
    __syncthreads();
    // coalesced write-out of dQ|dK|dV and the per-sequence column sums (bias gradients)
    const int ncol = hpb * dk;
    for (int third = 0; third < 3; ++third) {
        const float* src = smem + third * per;
        for (int idx = threadIdx.x; idx < L * ncol; idx += blockDim.x) {
            const int l = idx / ncol, c = idx - l * ncol;
            const int h2 = c / dk, d = c - h2 * dk;
            a.d_qkv[(row0 + l) * ld + third * D + h0 * dk + c] =
                src[(size_t)(h2 * L + l) * kRowStride + d];
        }
        for (int c = threadIdx.x; c < ncol; c += blockDim.x) {
            const int h2 = c / dk, d = c - h2 * dk;
            float sum = 0.f;
            for (int l = 0; l < L; ++l) sum += src[(size_t)(h2 * L + l) * kRowStride + d];
            a.d_bias_part[(long long)seq * ld + third * D + h0 * dk + c] = sum;
        }
    }
}

}  // namespace nrms
