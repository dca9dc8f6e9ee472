// attention_tile.cuh — register-tiled self-attention for sequences of up to 32 tokens (the
// title encoder: n_words_title = 20..30), forward and backward.  Same math and I/O contract as
// attention.cuh (reference nrms_v0.py:13-23, 46-76, 171-173), different mapping:
//
//   one WARP owns one (sequence, head); the 32x32 score tile lives in registers, each lane
//   holding a 4x8 sub-tile (rows ly+8i, columns lx+4j; lane = 4*ly + lx), so every product of
//   the head — S = Q K^T, O = P V and, in the backward, dP = dO V^T, dV = P^T dO, dK = dS^T Q,
//   dQ = dS K — is a small register-blocked GEMM fed by 16-byte shared-memory reads
//   (3 reads per 32 FMAs, all bank-conflict free with the 36-float row stride), and the row
//   softmax is two warp shuffles over the 4 lanes that share a row.
//
// Shared memory per head: three (forward) / four (backward) 32x36 fp32 slots, reused in place
// as operands die (P over Q, O over K; P over V, dS over dO, dV|dK|dQ over V|Q|K), so a CTA of
// 5 heads needs 69 KB / 92 KB and 2-3 CTAs share an SM.  Rows >= L and columns >= d_k of every
// slot are zero, which makes the padded 32x32x32 products exact.
#pragma once
#include "attention.cuh"

namespace nrms {

constexpr int kTile = 32;
constexpr int kSlot = kTile * kRowStride;   // floats per 32x36 slot

__host__ __device__ inline size_t attn_tile_fwd_smem_bytes(int hpb) {
    return (size_t)3 * hpb * kSlot * sizeof(float) + (size_t)kTile * 64;
}
__host__ __device__ inline size_t attn_tile_bwd_smem_bytes(int hpb) {
    return (size_t)4 * hpb * kSlot * sizeof(float);
}

// cp.async load of columns [col0, col0 + hpb*dk) of rows [row0, row0+L) into per-head 32x36
// slots dst[hh*kSlot + l*36 + d]; rows >= L and columns in [dk, 36) are zero-filled.
__device__ __forceinline__ void load_slots_async(float* dst, const float* src, long long row0, int ld,
                                                 int col0, int L, int dk, int hpb) {
    const int nu = (hpb * dk) >> 1;
    for (int u = threadIdx.x; u < nu; u += blockDim.x) {
        const int c = u << 1;
        const int hh = c / dk, d = c - hh * dk;
        const float* g = src + row0 * ld + col0 + c;
        float* t = dst + (size_t)hh * kSlot + d;
#pragma unroll 4
        for (int l = 0; l < L; ++l) cp_async8(t + l * kRowStride, g + (long long)l * ld);
    }
    const int npad = kRowStride - dk;
    for (int i = threadIdx.x; i < hpb * L * npad; i += blockDim.x) {
        const int r = i / npad, d = dk + (i - r * npad);           // r = hh*L + l
        const int hh = r / L, l = r - hh * L;
        dst[(size_t)hh * kSlot + l * kRowStride + d] = 0.f;
    }
    const int rpad = kTile - L;
    for (int i = threadIdx.x; i < hpb * rpad * kRowStride; i += blockDim.x) {
        const int hh = i / (rpad * kRowStride), o = i - hh * (rpad * kRowStride);
        dst[(size_t)hh * kSlot + L * kRowStride + o] = 0.f;
    }
}

// acc[i][j] += sum_k A[ly+8i][k] * B[lx+4j][k]     (both row-major slots, k = 0..31)
__device__ __forceinline__ void tile_abt(float (&acc)[4][8], const float* A, const float* B, int ly, int lx) {
#pragma unroll
    for (int kc = 0; kc < kTile / 4; ++kc) {
        float4 a[4], b[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(A + (ly + 8 * i) * kRowStride + 4 * kc);
#pragma unroll
        for (int j = 0; j < 8; ++j) b[j] = *reinterpret_cast<const float4*>(B + (lx + 4 * j) * kRowStride + 4 * kc);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                acc[i][j] = fmaf(a[i].x, b[j].x, acc[i][j]);
                acc[i][j] = fmaf(a[i].y, b[j].y, acc[i][j]);
                acc[i][j] = fmaf(a[i].z, b[j].z, acc[i][j]);
                acc[i][j] = fmaf(a[i].w, b[j].w, acc[i][j]);
            }
    }
}
// acc[i][d] += sum_k A[ly+8i][k] * B[k][8lx+d]     (k = 0..31)
__device__ __forceinline__ void tile_ab(float (&acc)[4][8], const float* A, const float* B, int ly, int lx) {
#pragma unroll
    for (int kc = 0; kc < kTile / 4; ++kc) {
        float4 a[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(A + (ly + 8 * i) * kRowStride + 4 * kc);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const float4 b0 = *reinterpret_cast<const float4*>(B + (4 * kc + t) * kRowStride + 8 * lx);
            const float4 b1 = *reinterpret_cast<const float4*>(B + (4 * kc + t) * kRowStride + 8 * lx + 4);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float av = t == 0 ? a[i].x : (t == 1 ? a[i].y : (t == 2 ? a[i].z : a[i].w));
                acc[i][0] = fmaf(av, b0.x, acc[i][0]); acc[i][1] = fmaf(av, b0.y, acc[i][1]);
                acc[i][2] = fmaf(av, b0.z, acc[i][2]); acc[i][3] = fmaf(av, b0.w, acc[i][3]);
                acc[i][4] = fmaf(av, b1.x, acc[i][4]); acc[i][5] = fmaf(av, b1.y, acc[i][5]);
                acc[i][6] = fmaf(av, b1.z, acc[i][6]); acc[i][7] = fmaf(av, b1.w, acc[i][7]);
            }
        }
    }
}
// acc[i][d] += sum_r A[r][4ly+i] * B[r][8lx+d]     (A^T B, r = 0..31; output rows 4ly+i)
__device__ __forceinline__ void tile_atb(float (&acc)[4][8], const float* A, const float* B, int ly, int lx) {
#pragma unroll 8
    for (int r = 0; r < kTile; ++r) {
        const float4 a = *reinterpret_cast<const float4*>(A + r * kRowStride + 4 * ly);
        const float4 b0 = *reinterpret_cast<const float4*>(B + r * kRowStride + 8 * lx);
        const float4 b1 = *reinterpret_cast<const float4*>(B + r * kRowStride + 8 * lx + 4);
        const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            acc[i][0] = fmaf(av[i], b0.x, acc[i][0]); acc[i][1] = fmaf(av[i], b0.y, acc[i][1]);
            acc[i][2] = fmaf(av[i], b0.z, acc[i][2]); acc[i][3] = fmaf(av[i], b0.w, acc[i][3]);
            acc[i][4] = fmaf(av[i], b1.x, acc[i][4]); acc[i][5] = fmaf(av[i], b1.y, acc[i][5]);
            acc[i][6] = fmaf(av[i], b1.z, acc[i][6]); acc[i][7] = fmaf(av[i], b1.w, acc[i][7]);
        }
    }
}
__device__ __forceinline__ void zero_tile(float (&acc)[4][8]) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
}
// dst[row_i][8lx + d] = acc[i][d] * mul[i]   (rows given by the caller's mapping)
__device__ __forceinline__ void store_tile(float* dst, const float (&acc)[4][8], const int (&rows)[4], int lx,
                                           const float (&mul)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float* p = dst + rows[i] * kRowStride + 8 * lx;
        *reinterpret_cast<float4*>(p) = make_float4(acc[i][0] * mul[i], acc[i][1] * mul[i], acc[i][2] * mul[i], acc[i][3] * mul[i]);
        *reinterpret_cast<float4*>(p + 4) = make_float4(acc[i][4] * mul[i], acc[i][5] * mul[i], acc[i][6] * mul[i], acc[i][7] * mul[i]);
    }
}
// scatter the score-layout tile: dst[ly+8i][lx+4j] = t[i][j]   (conflict-free scalar stores)
__device__ __forceinline__ void store_score_tile(float* dst, const float (&t)[4][8], int ly, int lx) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) dst[(ly + 8 * i) * kRowStride + lx + 4 * j] = t[i][j];
}
__device__ __forceinline__ float quad_max(float v) {
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
    return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

// ------------------------------------------------------------------------------------------------
// forward: grid (n_seq, ceil(n_heads/hpb)), block = hpb warps
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(320) attn_tile_fwd_kernel(const AttnArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int L = a.L, D = a.D, dk = a.dk;
    const int seq = blockIdx.x;
    const int h0 = blockIdx.y * a.hpb;
    const int hpb = min(a.hpb, a.n_heads - h0);
    float* Qs = smem;                                   // Q, later P
    float* Ks = Qs + (size_t)a.hpb * kSlot;             // K, later O (output staging)
    float* Vs = Ks + (size_t)a.hpb * kSlot;
    uint8_t* s_mask = reinterpret_cast<uint8_t*>(Vs + (size_t)a.hpb * kSlot);   // [32][64] keep bytes
    const long long row0 = (long long)seq * L;
    const int ld = 3 * D;
    const int col0 = h0 * dk, ncol = hpb * dk;

    load_slots_async(Qs, a.qkv, row0, ld, col0, L, dk, hpb);
    load_slots_async(Ks, a.qkv, row0, ld, D + col0, L, dk, hpb);
    load_slots_async(Vs, a.qkv, row0, ld, 2 * D + col0, L, dk, hpb);
    if (a.drop.enabled()) {
        const int g0 = col0 >> 3, g1 = (col0 + ncol + 7) >> 3;
        const int ng = g1 - g0;
        for (int i = threadIdx.x; i < L * ng; i += blockDim.x) {
            const int l = i / ng, g = g0 + (i - l * ng);
            const uint32_t keep = a.drop.keep8(kDropContext, (uint64_t)(row0 + l), (uint32_t)g);
            s_mask[l * 64 + g] = (uint8_t)keep;
            if (a.cmask && g < a.mask_bytes) a.cmask[(row0 + l) * a.mask_bytes + g] = (uint8_t)keep;
        }
    }
    cp_async_wait_all();
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ly = lane >> 2, lx = lane & 3;
    if (warp < hpb) {
        float* Qh = Qs + (size_t)warp * kSlot;
        float* Kh = Ks + (size_t)warp * kSlot;
        const float* Vh = Vs + (size_t)warp * kSlot;
        float s[4][8];
        zero_tile(s);
        tile_abt(s, Qh, Kh, ly, lx);
        // softmax over the keys: a row is shared by the 4 lanes of a quad
        float inv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float m = -INFINITY;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                s[i][j] = (lx + 4 * j < L) ? s[i][j] * a.scale : -INFINITY;   // scores / sqrt(d_k); no key >= L
                m = fmaxf(m, s[i][j]);
            }
            m = quad_max(m);
            float sum = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                s[i][j] = __expf(s[i][j] - m);
                sum += s[i][j];
            }
            sum = quad_sum(sum);
            inv[i] = 1.f / sum;
            const int r = ly + 8 * i;
            if (lx == 0 && r < L) a.lse[(row0 + r) * a.n_heads + h0 + warp] = m + __logf(sum);
        }
        __syncwarp();                       // every lane has finished reading Q
        store_score_tile(Qh, s, ly, lx);    // unnormalised P over Q
        __syncwarp();
        float o[4][8];
        zero_tile(o);
        tile_ab(o, Qh, Vh, ly, lx);
        __syncwarp();                       // (K was last read in tile_abt; P reads are done)
        const int rows[4] = {ly, ly + 8, ly + 16, ly + 24};
        store_tile(Kh, o, rows, lx, inv);   // O = P V / rowsum over K
    }
    __syncthreads();

    // coalesced write-out in 2-column units (+ context dropout, nrms_v0.py:171-173)
    const bool img = a.ctx_img.hi != nullptr;
    const bool drop = a.drop.enabled();
    for (int u = threadIdx.x; u < (ncol >> 1); u += blockDim.x) {
        const int c = u << 1;
        const int hh = c / dk, d = c - hh * dk;
        const int col = col0 + c;
        const float* srow = Ks + (size_t)hh * kSlot + d;
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
    if (img) {
        if (h0 + hpb == a.n_heads) {
            const int cpad = a.ctx_img.chunks * 64 - D;   // even
            for (int i = threadIdx.x; i < L * (cpad >> 1); i += blockDim.x) {
                const int l = i / (cpad >> 1), col = D + ((i - l * (cpad >> 1)) << 1);
                const long long off = ig::img_unit_off(a.ctx_img.chunk_stride, row0 + l, col >> 3) + (col & 7) * 2;
                *reinterpret_cast<uint32_t*>(a.ctx_img.hi + off) = 0u;
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
// backward: grid (n_seq, ceil(n_heads/hpb)), block = hpb warps
//   P = exp(scale*Q K^T - lse) ; dP = dO V^T ; dS = scale * P o (dP - delta) ; delta = rowsum(dO o O)
//   dV = P^T dO ; dK = dS^T Q ; dQ = dS K
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(320) attn_tile_bwd_kernel(const AttnArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int L = a.L, D = a.D, dk = a.dk;
    const int seq = blockIdx.x;
    const int h0 = blockIdx.y * a.hpb;
    const int hpb = min(a.hpb, a.n_heads - h0);
    const size_t per = (size_t)a.hpb * kSlot;
    float* Qs = smem;        // Q  -> dK
    float* Ks = Qs + per;    // K  -> dQ
    float* Vs = Ks + per;    // V  -> P -> dV
    float* Gs = Vs + per;    // dO -> dS
    const long long row0 = (long long)seq * L;
    const int ld = 3 * D;
    const int col0 = h0 * dk, ncol = hpb * dk;
    const bool drop = a.drop.enabled() && a.cmask != nullptr;

    load_slots_async(Qs, a.qkv, row0, ld, col0, L, dk, hpb);
    load_slots_async(Ks, a.qkv, row0, ld, D + col0, L, dk, hpb);
    load_slots_async(Vs, a.qkv, row0, ld, 2 * D + col0, L, dk, hpb);
    load_slots_async(Gs, a.d_ctx, row0, D, col0, L, dk, hpb);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ly = lane >> 2, lx = lane & 3;
    const bool has_head = warp < hpb;
    // post-dropout context of this lane's (row, 8 columns) pieces, straight from global while the
    // async copies fly: delta_i = sum_d d_ctx * ctx (both carry the same dropout factor)
    float2 ov[4][4];
    float lse[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = ly + 8 * i;
        lse[i] = (has_head && r < L) ? a.lse[(row0 + r) * a.n_heads + h0 + warp] : 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int d = 8 * lx + 2 * c;
            ov[i][c] = (has_head && r < L && d < dk)
                           ? __ldg(reinterpret_cast<const float2*>(a.ctx + (row0 + r) * D + col0 + warp * dk + d))
                           : make_float2(0.f, 0.f);
        }
    }
    cp_async_wait_all();
    __syncthreads();

    if (has_head) {
        float* Qh = Qs + (size_t)warp * kSlot;
        float* Kh = Ks + (size_t)warp * kSlot;
        float* Vh = Vs + (size_t)warp * kSlot;
        float* Gh = Gs + (size_t)warp * kSlot;
        float delta[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float* g = Gh + (ly + 8 * i) * kRowStride + 8 * lx;
            const float4 g0 = *reinterpret_cast<const float4*>(g), g1 = *reinterpret_cast<const float4*>(g + 4);
            float dl = g0.x * ov[i][0].x;
            dl = fmaf(g0.y, ov[i][0].y, dl); dl = fmaf(g0.z, ov[i][1].x, dl); dl = fmaf(g0.w, ov[i][1].y, dl);
            dl = fmaf(g1.x, ov[i][2].x, dl); dl = fmaf(g1.y, ov[i][2].y, dl);
            dl = fmaf(g1.z, ov[i][3].x, dl); dl = fmaf(g1.w, ov[i][3].y, dl);
            delta[i] = quad_sum(dl);
        }
        if (drop) {
            // dO = d_ctx * keep/(1-p), in place on this warp's own slot
            __syncwarp();
            for (int e = lane; e < L * (dk >> 1); e += 32) {
                const int l = e / (dk >> 1), d = (e - l * (dk >> 1)) << 1;
                const int col = col0 + warp * dk + d;
                const uint32_t keep = (uint32_t)a.cmask[(row0 + l) * a.mask_bytes + (col >> 3)] >> (col & 7);
                float2* p = reinterpret_cast<float2*>(Gh + l * kRowStride + d);
                float2 g = *p;
                g.x = (keep & 1u) ? g.x * a.drop.scale : 0.f;
                g.y = (keep & 2u) ? g.y * a.drop.scale : 0.f;
                *p = g;
            }
            __syncwarp();
        }
        // S and dP in the score layout
        float p[4][8], ds[4][8];
        zero_tile(p);
        zero_tile(ds);
        tile_abt(p, Qh, Kh, ly, lx);
        tile_abt(ds, Gh, Vh, ly, lx);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const bool row_ok = ly + 8 * i < L;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const bool ok = row_ok && (lx + 4 * j < L);
                const float pv = ok ? __expf(p[i][j] * a.scale - lse[i]) : 0.f;
                p[i][j] = pv;
                ds[i][j] = ok ? pv * (ds[i][j] - delta[i]) * a.scale : 0.f;
            }
        }
        __syncwarp();                        // all reads of V (dP) are done
        store_score_tile(Vh, p, ly, lx);     // P over V
        __syncwarp();
        const float one[4] = {1.f, 1.f, 1.f, 1.f};
        const int krows[4] = {4 * ly, 4 * ly + 1, 4 * ly + 2, 4 * ly + 3};
        float acc[4][8];
        zero_tile(acc);
        tile_atb(acc, Vh, Gh, ly, lx);       // dV[key][d] = sum_row P[row][key] dO[row][d]
        __syncwarp();                        // all reads of P and dO are done
        store_tile(Vh, acc, krows, lx, one); // dV over P
        store_score_tile(Gh, ds, ly, lx);    // dS over dO
        __syncwarp();
        zero_tile(acc);
        tile_atb(acc, Gh, Qh, ly, lx);       // dK[key][d] = sum_row dS[row][key] Q[row][d]
        float acc2[4][8];
        zero_tile(acc2);
        tile_ab(acc2, Gh, Kh, ly, lx);       // dQ[row][d] = sum_key dS[row][key] K[key][d]
        __syncwarp();                        // all reads of Q, K, dS are done
        store_tile(Qh, acc, krows, lx, one); // dK over Q
        const int rows[4] = {ly, ly + 8, ly + 16, ly + 24};
        store_tile(Kh, acc2, rows, lx, one); // dQ over K
    }
    __syncthreads();

    // write-out: third 0 (dQ) <- Ks, third 1 (dK) <- Qs, third 2 (dV) <- Vs ; bias partials
    const bool img = a.d_qkv_img.hi != nullptr;
    const int nu = ncol >> 1;
    for (int u = threadIdx.x; u < 3 * nu; u += blockDim.x) {
        const int third = u / nu;
        const int c = (u - third * nu) << 1;
        const int h2 = c / dk, d = c - h2 * dk;
        const int col = third * D + col0 + c;
        const float* base = third == 0 ? Ks : (third == 1 ? Qs : Vs);
        const float* srow = base + (size_t)h2 * kSlot + d;
        float sum0 = 0.f, sum1 = 0.f;
        for (int l = 0; l < L; ++l) {
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
        }
        a.d_bias_part[(long long)seq * ld + col] = sum0;
        a.d_bias_part[(long long)seq * ld + col + 1] = sum1;
    }
    if (img) {
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
