// attention.cuh — per-head scaled-dot-product self-attention over short sequences
// (title: 20-48 words, history: 50-200 news), forward and backward.
// Reference: ScaledDotProductAttention.forward nrms_v0.py:13-23 called from
// MultiHeadSelfAttention.forward nrms_v0.py:46-76 — softmax(Q K^T / sqrt(d_k)) V, NO mask
// (length=None always, nrms_v0.py:170,196), NO output projection; NewsEncoder then applies
// F.dropout to the context (nrms_v0.py:171-173).
//
// Layout: qkv [M, 3D] rows = tokens, [Q | K | V], head h owns columns h*dk .. h*dk+dk-1 of
// each third (the view/transpose of nrms_v0.py:53-58).  A CTA handles one sequence and a
// group of `hpb` heads staged in shared memory (rows padded to 36 floats: 16-byte aligned and
// conflict-free for lane-per-row float4 reads).  Scores never leave registers:
//   forward : a unit of U (16|32) lanes owns a head; each lane owns TWO query rows and streams
//             over the keys (broadcast float4 reads shared by both rows) with a softmax that
//             is rescaled once per group of 6 keys;
//   backward: one warp per (head, 32-row block); pass A: lane = query row (dQ), pass B:
//             lane = key row (dK, dV); P is recomputed from the saved log-sum-exp.
// Global traffic is coalesced 8-byte (two-column) units in both directions; the forward writes
// the context as fp32 (pooling input), as a split-bf16 image (additive-projection operand,
// gemm_img.cuh) and the dropout keep bits; the backward writes dQ|dK|dV as an image and/or fp32.
#pragma once
#include "common.cuh"
#include "gemm_img.cuh"

namespace nrms {

constexpr int kDkPad = 32;      // per-head dim padded to 32 (pad lanes are zero)
constexpr int kRowStride = 36;  // smem row stride in floats
constexpr int kKeyGroup = 6;    // keys per softmax rescale in the forward

struct AttnArgs {
    const float* qkv;    // [M, 3D]
    const uint16_t* qkv_hi;   // head-padded, head-blocked split-bf16 planes (attention_hp.cuh), instead of qkv
    const uint16_t* qkv_lo;
    float* ctx;          // fwd out / bwd in: post-dropout context [M, D]
    ig::Img ctx_img;     // fwd out (optional): image of the post-dropout context
    uint8_t* cmask;      // fwd out / bwd in (optional): keep bits [M, mask_bytes] (dropout only)
    float* lse;          // [M, n_heads] log-sum-exp of the scaled scores
    const float* d_ctx;  // bwd: grad wrt post-dropout context [M, D]
    float* d_qkv;        // bwd out (optional) fp32 [M, 3D]
    ig::Img d_qkv_img;   // bwd out (optional) image of [M, 3D]
    float* d_bias_part;  // bwd out (optional) [n_seq, 3D]: per-sequence column sums of d_qkv
    long long M;         // total token rows (n_seq * L)
    int L, D, n_heads, dk, hpb, mask_bytes;
    float scale;         // 1/sqrt(dk)
    Dropout drop;        // context dropout (stream kDropContext); disabled in eval / user encoder
};

__host__ __device__ inline size_t attn_fwd_smem_bytes(int L, int hpb) {
    return (size_t)3 * hpb * L * kRowStride * sizeof(float) + (size_t)L * 64;
}
__host__ __device__ inline size_t attn_bwd_smem_bytes(int L, int hpb) {
    return (size_t)4 * hpb * L * kRowStride * sizeof(float);
}

__device__ __forceinline__ void cp_async8(float* smem_dst, const float* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)),
                 "l"(gsrc)
                 : "memory");
}
__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)),
                 "l"(gsrc)
                 : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// Asynchronous (cp.async) load of `ncol` columns starting at `col0` of rows [row0, row0+L) of a
// row-major matrix with leading dimension ld into dst[(hh*L + l)*kRowStride + d] (column c ->
// head hh = c/dk, d = c%dk); pad lanes d in [dk, 32) are zeroed with plain stores.  A thread
// owns a 2-column unit (8 bytes; 1 column when !VEC2) and walks the rows, so the copies of a
// warp are coalesced, there is no per-element index arithmetic and nothing waits until
// cp_async_wait_all().
template <bool VEC2>
__device__ __forceinline__ void load_heads_async(float* dst, const float* src, long long row0, int ld,
                                                 int col0, int L, int dk, int hpb) {
    const int ncol = hpb * dk;
    const int nu = VEC2 ? ncol >> 1 : ncol;
    for (int u = threadIdx.x; u < nu; u += blockDim.x) {
        const int c = VEC2 ? u << 1 : u;
        const int hh = c / dk, d = c - hh * dk;
        const float* g = src + row0 * ld + col0 + c;
        float* t = dst + (size_t)hh * L * kRowStride + d;
#pragma unroll 4
        for (int l = 0; l < L; ++l) {
            if (VEC2) cp_async8(t + l * kRowStride, g + (long long)l * ld);
            else cp_async4(t + l * kRowStride, g + (long long)l * ld);
        }
    }
    const int npad = kDkPad - dk;
    if (npad > 0) {
        for (int i = threadIdx.x; i < hpb * L * npad; i += blockDim.x) {
            const int r = i / npad, d = dk + (i - r * npad);
            dst[r * kRowStride + d] = 0.f;
        }
    }
}

__device__ __forceinline__ void load_row32(float* r, const float* p) {
#pragma unroll
    for (int c = 0; c < kDkPad / 4; ++c) {
        const float4 v = *reinterpret_cast<const float4*>(p + 4 * c);
        r[4 * c] = v.x; r[4 * c + 1] = v.y; r[4 * c + 2] = v.z; r[4 * c + 3] = v.w;
    }
}
__device__ __forceinline__ void store_row32(float* p, const float* r, float mul) {
#pragma unroll
    for (int c = 0; c < kDkPad / 4; ++c)
        *reinterpret_cast<float4*>(p + 4 * c) =
            make_float4(r[4 * c] * mul, r[4 * c + 1] * mul, r[4 * c + 2] * mul, r[4 * c + 3] * mul);
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
template <int U, bool VEC2>
__global__ void __launch_bounds__(256) attn_fwd_kernel(const AttnArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int L = a.L, D = a.D, dk = a.dk;
    const int seq = blockIdx.x;
    const int h0 = blockIdx.y * a.hpb;
    const int hpb = min(a.hpb, a.n_heads - h0);
    const size_t per = (size_t)a.hpb * L * kRowStride;
    float* Qs = smem;   // Q, later the normalised output O
    float* Ks = Qs + per;
    float* Vs = Ks + per;
    uint8_t* s_mask = reinterpret_cast<uint8_t*>(Vs + per);   // [L][64] keep bytes (dropout only)
    const long long row0 = (long long)seq * L;
    const int ld = 3 * D;
    const int col0 = h0 * dk, ncol = hpb * dk;

    load_heads_async<VEC2>(Qs, a.qkv, row0, ld, col0, L, dk, hpb);
    load_heads_async<VEC2>(Ks, a.qkv, row0, ld, D + col0, L, dk, hpb);
    load_heads_async<VEC2>(Vs, a.qkv, row0, ld, 2 * D + col0, L, dk, hpb);
    if (a.drop.enabled()) {
        // keep bits of the 8-column groups overlapping this CTA's columns: one Philox call each
        const int g0 = col0 >> 3, g1 = (col0 + ncol + 7) >> 3;
        const int ng = g1 - g0;
        for (int i = threadIdx.x; i < L * ng; i += blockDim.x) {
            const int l = i / ng, g = g0 + (i - l * ng);
            const uint32_t keep = a.drop.keep8(kDropContext, (uint64_t)(row0 + l), (uint32_t)g);
            s_mask[l * 64 + g] = (uint8_t)keep;
            // a group straddling two head groups is written twice with the same value
            if (a.cmask && g < a.mask_bytes) a.cmask[(row0 + l) * a.mask_bytes + g] = (uint8_t)keep;
        }
    }
    cp_async_wait_all();
    __syncthreads();

    const int unit = threadIdx.x / U, ul = threadIdx.x % U;
    const int units = blockDim.x / U;
    const int rb_n = ceil_div(L, 2 * U);
    for (int task = unit; task < hpb * rb_n; task += units) {
        const int hh = task / rb_n;
        const int r0 = (task - hh * rb_n) * 2 * U + ul, r1 = r0 + U;
        const bool act0 = r0 < L, act1 = r1 < L;
        float q0[kDkPad], q1[kDkPad], acc0[kDkPad], acc1[kDkPad];
        load_row32(q0, Qs + (size_t)(hh * L + (act0 ? r0 : 0)) * kRowStride);
        load_row32(q1, Qs + (size_t)(hh * L + (act1 ? r1 : 0)) * kRowStride);
#pragma unroll
        for (int d = 0; d < kDkPad; ++d) {
            q0[d] *= a.scale; q1[d] *= a.scale;   // scores = (Q K^T) / sqrt(d_k)  (nrms_v0.py:14)
            acc0[d] = 0.f; acc1[d] = 0.f;
        }
        float m0 = -INFINITY, m1 = -INFINITY, den0 = 0.f, den1 = 0.f;
        const float* kbase = Ks + (size_t)hh * L * kRowStride;
        const float* vbase = Vs + (size_t)hh * L * kRowStride;
        for (int j0 = 0; j0 < L; j0 += kKeyGroup) {
            float s0[kKeyGroup], s1[kKeyGroup];
#pragma unroll
            for (int g = 0; g < kKeyGroup; ++g) {
                const int j = min(j0 + g, L - 1);
                const float* kr = kbase + j * kRowStride;
                float x0 = 0.f, x1 = 0.f;
#pragma unroll
                for (int c = 0; c < kDkPad / 4; ++c) {
                    const float4 kv = *reinterpret_cast<const float4*>(kr + 4 * c);
                    x0 = fmaf(q0[4 * c], kv.x, x0); x1 = fmaf(q1[4 * c], kv.x, x1);
                    x0 = fmaf(q0[4 * c + 1], kv.y, x0); x1 = fmaf(q1[4 * c + 1], kv.y, x1);
                    x0 = fmaf(q0[4 * c + 2], kv.z, x0); x1 = fmaf(q1[4 * c + 2], kv.z, x1);
                    x0 = fmaf(q0[4 * c + 3], kv.w, x0); x1 = fmaf(q1[4 * c + 3], kv.w, x1);
                }
                const bool valid = j0 + g < L;
                s0[g] = valid ? x0 : -INFINITY;
                s1[g] = valid ? x1 : -INFINITY;
            }
            float n0 = m0, n1 = m1;
#pragma unroll
            for (int g = 0; g < kKeyGroup; ++g) { n0 = fmaxf(n0, s0[g]); n1 = fmaxf(n1, s1[g]); }
            const float c0 = __expf(m0 - n0), c1 = __expf(m1 - n1);   // exp(-inf) = 0 on the first group
            m0 = n0; m1 = n1;
            den0 *= c0; den1 *= c1;
#pragma unroll
            for (int d = 0; d < kDkPad; ++d) { acc0[d] *= c0; acc1[d] *= c1; }
#pragma unroll
            for (int g = 0; g < kKeyGroup; ++g) {
                const int j = min(j0 + g, L - 1);
                const float p0 = __expf(s0[g] - n0), p1 = __expf(s1[g] - n1);   // 0 for invalid keys
                den0 += p0; den1 += p1;
                const float* vr = vbase + j * kRowStride;
#pragma unroll
                for (int c = 0; c < kDkPad / 4; ++c) {
                    const float4 vv = *reinterpret_cast<const float4*>(vr + 4 * c);
                    acc0[4 * c] = fmaf(p0, vv.x, acc0[4 * c]); acc1[4 * c] = fmaf(p1, vv.x, acc1[4 * c]);
                    acc0[4 * c + 1] = fmaf(p0, vv.y, acc0[4 * c + 1]); acc1[4 * c + 1] = fmaf(p1, vv.y, acc1[4 * c + 1]);
                    acc0[4 * c + 2] = fmaf(p0, vv.z, acc0[4 * c + 2]); acc1[4 * c + 2] = fmaf(p1, vv.z, acc1[4 * c + 2]);
                    acc0[4 * c + 3] = fmaf(p0, vv.w, acc0[4 * c + 3]); acc1[4 * c + 3] = fmaf(p1, vv.w, acc1[4 * c + 3]);
                }
            }
        }
        // own rows of Qs: nobody else reads them, safe to overwrite with the output
        if (act0) {
            store_row32(Qs + (size_t)(hh * L + r0) * kRowStride, acc0, 1.f / den0);
            a.lse[(row0 + r0) * a.n_heads + h0 + hh] = m0 + __logf(den0);
        }
        if (act1) {
            store_row32(Qs + (size_t)(hh * L + r1) * kRowStride, acc1, 1.f / den1);
            a.lse[(row0 + r1) * a.n_heads + h0 + hh] = m1 + __logf(den1);
        }
    }
    __syncthreads();

    // coalesced write-out in 2-column units (+ context dropout, nrms_v0.py:171-173)
    const bool img = a.ctx_img.hi != nullptr;
    const bool drop = a.drop.enabled();
    if (VEC2) {
        for (int u = threadIdx.x; u < (ncol >> 1); u += blockDim.x) {
            const int c = u << 1;
            const int hh = c / dk, d = c - hh * dk;
            const int col = col0 + c;
            const float* srow = Qs + (size_t)hh * L * kRowStride + d;
            for (int l = 0; l < L; ++l) {
                float2 v = *reinterpret_cast<const float2*>(srow + l * kRowStride);
                if (drop) {
                    const uint32_t keep = (uint32_t)s_mask[l * 64 + (col >> 3)] >> (col & 7);
                    v.x = (keep & 1u) ? v.x * a.drop.scale : 0.f;
                    v.y = (keep & 2u) ? v.y * a.drop.scale : 0.f;
                }
                *reinterpret_cast<float2*>(a.ctx + (row0 + l) * D + col) = v;
                if (img) {
                    __nv_bfloat16 h0b, l0b, h1b, l1b;
                    tc::split_bf16(v.x, h0b, l0b);
                    tc::split_bf16(v.y, h1b, l1b);
                    const long long off = ig::img_unit_off(a.ctx_img.chunk_stride, row0 + l, col >> 3) + (col & 7) * 2;
                    *reinterpret_cast<uint32_t*>(a.ctx_img.hi + off) =
                        (uint32_t)__bfloat16_as_ushort(h0b) | ((uint32_t)__bfloat16_as_ushort(h1b) << 16);
                    *reinterpret_cast<uint32_t*>(a.ctx_img.lo + off) =
                        (uint32_t)__bfloat16_as_ushort(l0b) | ((uint32_t)__bfloat16_as_ushort(l1b) << 16);
                }
            }
        }
    } else {
        for (int c = threadIdx.x; c < ncol; c += blockDim.x) {
            const int hh = c / dk, d = c - hh * dk;
            const int col = col0 + c;
            for (int l = 0; l < L; ++l) {
                float v = Qs[(size_t)(hh * L + l) * kRowStride + d];
                if (drop) v = ((s_mask[l * 64 + (col >> 3)] >> (col & 7)) & 1u) ? v * a.drop.scale : 0.f;
                a.ctx[(row0 + l) * D + col] = v;
            }
        }
    }
    if (img) {
        // image padding: columns [D, 64*chunks) of this sequence's rows (the last head group does
        // it) and, by the very last CTA, the rows [M, rows_pad) — both enter GEMM reductions
        if (h0 + hpb == a.n_heads) {
            const int cpad = a.ctx_img.chunks * 64 - D;   // even
            for (int i = threadIdx.x; i < L * (cpad >> 1); i += blockDim.x) {
                const int l = i / (cpad >> 1), col = D + ((i - l * (cpad >> 1)) << 1);
                const long long off = ig::img_unit_off(a.ctx_img.chunk_stride, row0 + l, col >> 3) + (col & 7) * 2;
                // column D = 1.0 (bf16 {1.0, 0.0}): the weight-gradient GEMM's bias column (gather.cuh)
                *reinterpret_cast<uint32_t*>(a.ctx_img.hi + off) = col == D ? 0x00003F80u : 0u;
                *reinterpret_cast<uint32_t*>(a.ctx_img.lo + off) = 0u;
            }
        }
        if (seq == (int)gridDim.x - 1 && blockIdx.y == 0) {
            const int groups = a.ctx_img.chunks * 8;
            const long long npad = a.ctx_img.rows_pad - a.M;
            for (long long i = threadIdx.x; i < npad * groups; i += blockDim.x)
                ig::img_store8_zero(a.ctx_img, a.M + i / groups, (int)(i % groups));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// backward.  With P = softmax(S), S = scale * Q K^T, O = P V, dO given:
//   dV = P^T dO ; dP = dO V^T ; dS = P * (dP - rowsum(dO*O)) ; dQ = scale dS K ; dK = scale dS^T Q
// ------------------------------------------------------------------------------------------------
template <bool VEC2>
__global__ void __launch_bounds__(256) attn_bwd_kernel(const AttnArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int L = a.L, D = a.D, dk = a.dk;
    const int seq = blockIdx.x;
    const int h0 = blockIdx.y * a.hpb;
    const int hpb = min(a.hpb, a.n_heads - h0);
    const size_t per = (size_t)a.hpb * L * kRowStride;
    float* Qs = smem;  // Q (unscaled)
    float* Ks = Qs + per;
    float* Vs = Ks + per;
    float* Gs = Vs + per;  // dO (grad wrt pre-dropout context); spare columns 32,33: delta, lse
    const long long row0 = (long long)seq * L;
    const int ld = 3 * D;
    const int col0 = h0 * dk, ncol = hpb * dk;
    const bool drop = a.drop.enabled() && a.cmask != nullptr;

    // phase 0: dO = d_ctx * mask -> Gs, post-dropout context -> Qs (temporarily), then
    // delta_i = sum_d dO_raw * O_raw = (sum_d dO * ctx_post) / drop.scale, one (head,row) per
    // thread in a fixed order (deterministic; no atomics)
    load_heads_async<VEC2>(Ks, a.qkv, row0, ld, D + col0, L, dk, hpb);
    load_heads_async<VEC2>(Vs, a.qkv, row0, ld, 2 * D + col0, L, dk, hpb);
    load_heads_async<VEC2>(Qs, a.ctx, row0, D, col0, L, dk, hpb);
    load_heads_async<VEC2>(Gs, a.d_ctx, row0, D, col0, L, dk, hpb);
    cp_async_wait_all();
    __syncthreads();
    const float inv_drop = drop ? 1.f / a.drop.scale : 1.f;
    for (int i = threadIdx.x; i < hpb * L; i += blockDim.x) {
        const int hh = i / L, l = i - hh * L;
        float* g = Gs + (size_t)i * kRowStride;
        const float* o = Qs + (size_t)i * kRowStride;
        if (drop) {
            for (int d = 0; d < dk; ++d) {
                const int col = col0 + hh * dk + d;
                const uint32_t keep = ((uint32_t)a.cmask[(row0 + l) * a.mask_bytes + (col >> 3)] >> (col & 7)) & 1u;
                g[d] = keep ? g[d] * a.drop.scale : 0.f;
            }
        }
        float dl = 0.f;
#pragma unroll
        for (int c = 0; c < kDkPad / 4; ++c) {
            const float4 gv = *reinterpret_cast<const float4*>(g + 4 * c);
            const float4 ov = *reinterpret_cast<const float4*>(o + 4 * c);
            dl = fmaf(gv.x, ov.x, dl); dl = fmaf(gv.y, ov.y, dl);
            dl = fmaf(gv.z, ov.z, dl); dl = fmaf(gv.w, ov.w, dl);
        }
        g[kDkPad] = dl * inv_drop;
        g[kDkPad + 1] = a.lse[(row0 + l) * a.n_heads + h0 + hh];
    }
    __syncthreads();
    load_heads_async<VEC2>(Qs, a.qkv, row0, ld, col0, L, dk, hpb);
    cp_async_wait_all();
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rc = ceil_div(L, 32);
    // one (head, 32-row block) task per warp (the launch configuration guarantees it); results
    // stay in registers until every warp has finished reading shared memory
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
        load_row32(q, qb + ii * kRowStride);
        load_row32(go, gb + ii * kRowStride);
#pragma unroll
        for (int d = 0; d < kDkPad; ++d) q[d] *= a.scale;
        const float delta = gb[ii * kRowStride + kDkPad];
        const float lse = gb[ii * kRowStride + kDkPad + 1];
#pragma unroll
        for (int d = 0; d < kDkPad; ++d) dq[d] = 0.f;
        if (has_task) {
#pragma unroll 1
            for (int j = 0; j < L; ++j) {
                const float* kr = kb + j * kRowStride;
                const float* vr = vb + j * kRowStride;
                // four partial sums per dot product: dependent FMA chains of 8, not 32
                float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f), d4 = s4;
                float kreg[kDkPad];
#pragma unroll
                for (int c = 0; c < kDkPad / 4; ++c) {
                    const float4 kv = *reinterpret_cast<const float4*>(kr + 4 * c);
                    const float4 vv = *reinterpret_cast<const float4*>(vr + 4 * c);
                    kreg[4 * c] = kv.x; kreg[4 * c + 1] = kv.y; kreg[4 * c + 2] = kv.z;
                    kreg[4 * c + 3] = kv.w;
                    s4.x = fmaf(q[4 * c], kv.x, s4.x); s4.y = fmaf(q[4 * c + 1], kv.y, s4.y);
                    s4.z = fmaf(q[4 * c + 2], kv.z, s4.z); s4.w = fmaf(q[4 * c + 3], kv.w, s4.w);
                    d4.x = fmaf(go[4 * c], vv.x, d4.x); d4.y = fmaf(go[4 * c + 1], vv.y, d4.y);
                    d4.z = fmaf(go[4 * c + 2], vv.z, d4.z); d4.w = fmaf(go[4 * c + 3], vv.w, d4.w);
                }
                const float s = (s4.x + s4.y) + (s4.z + s4.w), dp = (d4.x + d4.y) + (d4.z + d4.w);
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
        load_row32(k, kb + ii * kRowStride);
        load_row32(v, vb + ii * kRowStride);
#pragma unroll
        for (int d = 0; d < kDkPad; ++d) { dkk[d] = 0.f; dvv[d] = 0.f; }
        if (has_task) {
#pragma unroll 1
            for (int r = 0; r < L; ++r) {
                const float* qr = qb + r * kRowStride;  // Q_r (unscaled)
                const float* gr = gb + r * kRowStride;
                float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f), d4 = s4;
                float qreg[kDkPad], greg[kDkPad];
#pragma unroll
                for (int c = 0; c < kDkPad / 4; ++c) {
                    const float4 qv = *reinterpret_cast<const float4*>(qr + 4 * c);
                    const float4 gv = *reinterpret_cast<const float4*>(gr + 4 * c);
                    qreg[4 * c] = qv.x; qreg[4 * c + 1] = qv.y; qreg[4 * c + 2] = qv.z;
                    qreg[4 * c + 3] = qv.w;
                    greg[4 * c] = gv.x; greg[4 * c + 1] = gv.y; greg[4 * c + 2] = gv.z;
                    greg[4 * c + 3] = gv.w;
                    s4.x = fmaf(qv.x, k[4 * c], s4.x); s4.y = fmaf(qv.y, k[4 * c + 1], s4.y);
                    s4.z = fmaf(qv.z, k[4 * c + 2], s4.z); s4.w = fmaf(qv.w, k[4 * c + 3], s4.w);
                    d4.x = fmaf(gv.x, v[4 * c], d4.x); d4.y = fmaf(gv.y, v[4 * c + 1], d4.y);
                    d4.z = fmaf(gv.z, v[4 * c + 2], d4.z); d4.w = fmaf(gv.w, v[4 * c + 3], d4.w);
                }
                const float s = (s4.x + s4.y) + (s4.z + s4.w), dp = (d4.x + d4.y) + (d4.z + d4.w);
                const float delta = gr[kDkPad], lse = gr[kDkPad + 1];
                const float p = __expf(s * a.scale - lse);
                const float ds = p * (dp - delta) * a.scale;
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
        store_row32(Qs + (size_t)(hh * L + i) * kRowStride, dq, 1.f);
        store_row32(Ks + (size_t)(hh * L + i) * kRowStride, dkk, 1.f);
        store_row32(Vs + (size_t)(hh * L + i) * kRowStride, dvv, 1.f);
    }
    __syncthreads();
    // write-out of dQ|dK|dV and the per-sequence column sums (bias gradients)
    const bool img = a.d_qkv_img.hi != nullptr;
    // a thread owns a 2-column unit of one third and walks the rows: coalesced 8-byte stores, the
    // bias partial (column sum over the sequence) falls out of the same walk in a fixed order
    const int nu = VEC2 ? ncol >> 1 : ncol;
    for (int u = threadIdx.x; u < 3 * nu; u += blockDim.x) {
        const int third = u / nu;
        const int c = VEC2 ? (u - third * nu) << 1 : (u - third * nu);
        const int h2 = c / dk, d = c - h2 * dk;
        const int col = third * D + col0 + c;
        const float* srow = smem + third * per + (size_t)h2 * L * kRowStride + d;
        float sum0 = 0.f, sum1 = 0.f;
        for (int l = 0; l < L; ++l) {
            if (VEC2) {
                const float2 v = *reinterpret_cast<const float2*>(srow + l * kRowStride);
                sum0 += v.x; sum1 += v.y;
                if (a.d_qkv) *reinterpret_cast<float2*>(a.d_qkv + (row0 + l) * ld + col) = v;
                if (img) {
                    __nv_bfloat16 h0b, l0b, h1b, l1b;
                    tc::split_bf16(v.x, h0b, l0b);
                    tc::split_bf16(v.y, h1b, l1b);
                    const long long off = ig::img_unit_off(a.d_qkv_img.chunk_stride, row0 + l, col >> 3) + (col & 7) * 2;
                    *reinterpret_cast<uint32_t*>(a.d_qkv_img.hi + off) =
                        (uint32_t)__bfloat16_as_ushort(h0b) | ((uint32_t)__bfloat16_as_ushort(h1b) << 16);
                    *reinterpret_cast<uint32_t*>(a.d_qkv_img.lo + off) =
                        (uint32_t)__bfloat16_as_ushort(l0b) | ((uint32_t)__bfloat16_as_ushort(l1b) << 16);
                }
            } else {
                const float v = srow[l * kRowStride];
                sum0 += v;
                if (a.d_qkv) a.d_qkv[(row0 + l) * ld + col] = v;
            }
        }
        if (a.d_bias_part) {
            a.d_bias_part[(long long)seq * ld + col] = sum0;
            if (VEC2) a.d_bias_part[(long long)seq * ld + col + 1] = sum1;
        }
    }
    if (img) {
        // image padding: columns [3D, 16*ceil(3D/16)) are read by the data-gradient GEMM's last
        // k-step, rows [M, rows_pad) by the weight-gradient reduction: both must be zero
        if (h0 + hpb == a.n_heads) {
            const int cend = ceil_div(3 * D, 16) * 16;
            const int cpad = cend - 3 * D;   // even
            for (int idx = threadIdx.x; idx < L * (cpad >> 1); idx += blockDim.x) {
                const int l = idx / (cpad >> 1), col = 3 * D + ((idx - l * (cpad >> 1)) << 1);
                const long long off = ig::img_unit_off(a.d_qkv_img.chunk_stride, row0 + l, col >> 3) + (col & 7) * 2;
                *reinterpret_cast<uint32_t*>(a.d_qkv_img.hi + off) = 0u;
                *reinterpret_cast<uint32_t*>(a.d_qkv_img.lo + off) = 0u;
            }
        }
        if (seq == (int)gridDim.x - 1 && blockIdx.y == 0) {
            const int groups = a.d_qkv_img.chunks * 8;
            const long long npad = a.d_qkv_img.rows_pad - a.M;
            for (long long idx = threadIdx.x; idx < npad * groups; idx += blockDim.x)
                ig::img_store8_zero(a.d_qkv_img, a.M + idx / groups, (int)(idx % groups));
        }
    }
}

}  // namespace nrms
