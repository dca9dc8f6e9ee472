// attention_tile.cuh — register-tiled self-attention for sequences of up to 32 tokens (the
// title encoder: n_words_title = 20..30), forward and backward.  Same math and I/O contract as
// attention.cuh (reference nrms_v0.py:13-23, 46-76, 171-173), different mapping:
//
//   one WARP owns one (sequence, head) and never synchronises with another warp.  The 32x32
//   score tile lives in registers, each lane holding a 4x8 sub-tile (rows ly+8i, columns
//   lx+4j; lane = 4*ly + lx), so every product of the head — S = Q K^T, O = P V and, in the
//   backward, dP = dO V^T, dV = P^T dO, dK = dS^T Q, dQ = dS K — is a small register-blocked
//   GEMM fed by 16-byte shared-memory reads (3 reads per 32 FMAs, bank-conflict free with the
//   36-float row stride), and the row softmax is two warp shuffles over the 4 lanes of a row.
//
// Per warp: its head's Q/K/V (and dO) arrive by cp.async into private 32x36 fp32 slots (rows
// >= L and columns >= d_k zeroed, which makes the padded 32x32x32 products exact); P and dS
// reuse dead slots (P over Q; P over V, dS over dO); results leave straight from registers:
// fp32 context / gradient, split-bf16 image units (gemm_img.cuh), dropout keep bits, and — in the
// backward — the per-sequence bias partials by a shuffle reduction over the rows.
#pragma once
#include "attention.cuh"

namespace nrms {

constexpr int kTile = 32;
constexpr int kSlot = kTile * kRowStride;   // floats per 32x36 slot
constexpr int kTileWarps = 4;               // warps (= sequence-heads) per CTA

__host__ __device__ inline size_t attn_tile_fwd_smem_bytes() {
    return (size_t)kTileWarps * (3 * kSlot * sizeof(float) + kTile * 8);
}
__host__ __device__ inline size_t attn_tile_bwd_smem_bytes() {
    return (size_t)kTileWarps * 4 * kSlot * sizeof(float);
}

// Warp-private cp.async load of columns [col, col+dk) of rows [row0, row0+L) into one 32x36
// slot; rows >= L and columns in [dk, 36) are zero-filled with plain stores.  16 lanes walk the
// even rows, 16 the odd rows, each lane owning one 2-column unit.
template <int RS = kRowStride>
__device__ __forceinline__ void load_slot_async(float* slot, const float* src, long long row0, int ld,
                                                int col, int L, int dk, int lane) {
    const int pp = lane & 15, par = lane >> 4;
    if (2 * pp < dk) {
        const float* g = src + (row0 + par) * ld + col + 2 * pp;
        uint32_t t = (uint32_t)__cvta_generic_to_shared(slot + par * RS + 2 * pp);
        for (int l = par; l < L; l += 2) {
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(t), "l"(g) : "memory");
            t += 2 * RS * 4;
            g += 2 * ld;
        }
    }
    {
        float* row = slot + lane * RS;      // lane = slot row
        if (lane < L) {
            for (int d = dk; d < RS; d += 2) *reinterpret_cast<float2*>(row + d) = make_float2(0.f, 0.f);
        } else {
#pragma unroll
            for (int d = 0; d < RS; d += 4) *reinterpret_cast<float4*>(row + d) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
}

// Warp-level write-out of a finished 32x36 slot (rows [0,L), columns [0,dk)) to global column
// gcol0 + c of an fp32 matrix (may be null) and/or a split-bf16 image; optional dropout keep
// bits smask[row*8 + (col>>3) - g0]; SUMS adds the per-sequence column sums (bias partials).
// 16 lanes per row (one 2-column unit each), two rows per pass of a ROLLED loop: the code stays
// small (the kernels are instruction-fetch sensitive) and every store is coalesced.
template <bool SUMS, int RS = kRowStride>
__device__ __forceinline__ void warp_write_slot(const float* slot, int L, int dk, long long row0, int gcol0, float* out,
                                                int ld, const ig::Img& img, const uint8_t* smask, int g0,
                                                float drop_scale, float* sums, int lane) {
    const int c = (lane & 15) << 1, par = lane >> 4;
    const bool active = c < dk;
    const int col = gcol0 + c, g = col >> 3, gu = g & 7;
    const bool has_img = img.hi != nullptr;
    const long long choff = has_img ? (long long)(g >> 3) * img.chunk_stride + (col & 7) * 2 : 0;
#pragma unroll 1
    for (int l = par; l < L; l += 2) {
        if (!active) continue;
        float2 v = *reinterpret_cast<const float2*>(slot + l * RS + c);
        if (smask) {
            const uint32_t keep = (uint32_t)smask[l * 8 + g - g0] >> (col & 7);
            v.x = (keep & 1u) ? v.x * drop_scale : 0.f;
            v.y = (keep & 2u) ? v.y * drop_scale : 0.f;
        }
        const long long r = row0 + l;
        if (out) *reinterpret_cast<float2*>(out + r * ld + col) = v;
        if (has_img) {
            __nv_bfloat16 h0b, l0b, h1b, l1b;
            tc::split_bf16(v.x, h0b, l0b);
            tc::split_bf16(v.y, h1b, l1b);
            const int r7 = (int)(r & 7);
            const long long off = choff + (r >> 3) * 1024 + r7 * 128 + ((gu ^ r7) << 4);
            *reinterpret_cast<uint32_t*>(img.hi + off) =
                (uint32_t)__bfloat16_as_ushort(h0b) | ((uint32_t)__bfloat16_as_ushort(h1b) << 16);
            *reinterpret_cast<uint32_t*>(img.lo + off) =
                (uint32_t)__bfloat16_as_ushort(l0b) | ((uint32_t)__bfloat16_as_ushort(l1b) << 16);
        }
    }
    if (SUMS && sums != nullptr && lane < dk) {
        float s0 = 0.f, s1 = 0.f;
#pragma unroll 1
        for (int l = 0; l + 1 < L; l += 2) {
            s0 += slot[l * RS + lane];
            s1 += slot[(l + 1) * RS + lane];
        }
        if (L & 1) s0 += slot[(L - 1) * RS + lane];
        sums[gcol0 + lane] = s0 + s1;
    }
}
// padding of an image the kernel fills: columns [c0, c1) of this sequence's rows by the warp of
// the last head (zero; with `ones_col` column c0 is 1.0 so that the weight-gradient GEMM yields
// the bias gradient as its column c0), rows [M, rows_pad) by the very last warp of the grid
__device__ __forceinline__ void pad_image(const ig::Img& img, long long row0, int L, int c0, int c1, bool last_head,
                                          bool last_item, long long M, int lane, bool ones_col = false) {
    if (last_head) {
        const int np = (c1 - c0) >> 1;
        for (int i = lane; i < L * np; i += 32) {
            const int l = i / np, col = c0 + ((i - l * np) << 1);
            const long long off = ig::img_unit_off(img.chunk_stride, row0 + l, col >> 3) + (col & 7) * 2;
            *reinterpret_cast<uint32_t*>(img.hi + off) = (ones_col && col == c0) ? 0x00003F80u : 0u;   // bf16 {1.0, 0.0}
            *reinterpret_cast<uint32_t*>(img.lo + off) = 0u;
        }
    }
    if (last_item) {
        const int groups = img.chunks * 8;
        const long long npad = img.rows_pad - M;
        for (long long i = lane; i < npad * groups; i += 32) ig::img_store8_zero(img, M + i / groups, (int)(i % groups));
    }
}

// acc[i][j] += sum_k A[ly+8i][k] * B[lx+4j][k]     (both row-major slots, k = 0..31)
__device__ __forceinline__ void tile_abt(float (&acc)[4][8], const float* A, const float* B, int ly, int lx) {
    // partially unrolled on purpose: fully unrolled, the kernels were 120-180 KB of SASS and the
    // independent warps stalled on instruction fetch (no_inst 20-30% of samples)
#pragma unroll 2
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
#pragma unroll 2
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
// forward: one warp per (sequence, head); grid = ceil(n_seq*n_heads / kTileWarps)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTileWarps * 32) attn_tile_fwd_kernel(const AttnArgs a, long long n_items) {
    extern __shared__ __align__(16) float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long item = (long long)blockIdx.x * kTileWarps + warp;
    if (item >= n_items) return;
    const int L = a.L, D = a.D, dk = a.dk;
    const long long seq = item / a.n_heads;
    const int h = (int)(item - seq * a.n_heads);
    float* Qh = smem + (size_t)warp * 3 * kSlot;      // Q, later P
    float* Kh = Qh + kSlot;
    float* Vh = Kh + kSlot;
    uint8_t* smask = reinterpret_cast<uint8_t*>(smem + (size_t)kTileWarps * 3 * kSlot) + warp * kTile * 8;
    const long long row0 = seq * L;
    const int ld = 3 * D, col = h * dk;
    const int ly = lane >> 2, lx = lane & 3;

    load_slot_async(Qh, a.qkv, row0, ld, col, L, dk, lane);
    load_slot_async(Kh, a.qkv, row0, ld, D + col, L, dk, lane);
    load_slot_async(Vh, a.qkv, row0, ld, 2 * D + col, L, dk, lane);
    const int g0 = col >> 3;
    const bool drop = a.drop.enabled();
    if (drop) {
        // keep bits of the (up to 5) 8-column groups this head overlaps: one Philox call each; a
        // group shared with the neighbouring head is computed by both warps with identical bits
        const int ng = ((col + dk + 7) >> 3) - g0;
        for (int it = lane; it < L * 8; it += 32) {
            const int l = it >> 3, gi = it & 7;
            if (gi < ng) {
                const uint32_t keep = a.drop.keep8(kDropContext, (uint64_t)(row0 + l), (uint32_t)(g0 + gi));
                smask[l * 8 + gi] = (uint8_t)keep;
                if (a.cmask && g0 + gi < a.mask_bytes) a.cmask[(row0 + l) * a.mask_bytes + g0 + gi] = (uint8_t)keep;
            }
        }
    }
    cp_async_wait_all();
    __syncwarp();

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
        if (lx == 0 && r < L) a.lse[(row0 + r) * a.n_heads + h] = m + __logf(sum);
    }
    __syncwarp();                       // every lane has finished reading Q
    store_score_tile(Qh, s, ly, lx);    // unnormalised P over Q
    __syncwarp();
    float o[4][8];
    zero_tile(o);
    tile_ab(o, Qh, Vh, ly, lx);
    // O = P V / rowsum over the dead K slot, then (+ context dropout, nrms_v0.py:171-173) out
    const int rows[4] = {ly, ly + 8, ly + 16, ly + 24};
    store_tile(Kh, o, rows, lx, inv);
    __syncwarp();
    warp_write_slot<false>(Kh, L, dk, row0, col, a.ctx, D, a.ctx_img, drop ? smask : nullptr, g0, a.drop.scale, nullptr,
                           lane);
    if (a.ctx_img.hi != nullptr)
        pad_image(a.ctx_img, row0, L, D, a.ctx_img.chunks * 64, h == a.n_heads - 1, item == n_items - 1, a.M, lane);
}

// ------------------------------------------------------------------------------------------------
// backward: one warp per (sequence, head)
//   P = exp(scale*Q K^T - lse) ; dP = dO V^T ; dS = scale * P o (dP - delta) ; delta = rowsum(dO o O)
//   dV = P^T dO ; dK = dS^T Q ; dQ = dS K
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTileWarps * 32) attn_tile_bwd_kernel(const AttnArgs a, long long n_items) {
    extern __shared__ __align__(16) float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long item = (long long)blockIdx.x * kTileWarps + warp;
    if (item >= n_items) return;
    const int L = a.L, D = a.D, dk = a.dk;
    const long long seq = item / a.n_heads;
    const int h = (int)(item - seq * a.n_heads);
    float* Qh = smem + (size_t)warp * 4 * kSlot;
    float* Kh = Qh + kSlot;
    float* Vh = Kh + kSlot;    // V, later P
    float* Gh = Vh + kSlot;    // dO, later dS
    const long long row0 = seq * L;
    const int ld = 3 * D, col = h * dk;
    const int ly = lane >> 2, lx = lane & 3;
    const bool drop = a.drop.enabled() && a.cmask != nullptr;

    load_slot_async(Qh, a.qkv, row0, ld, col, L, dk, lane);
    load_slot_async(Kh, a.qkv, row0, ld, D + col, L, dk, lane);
    load_slot_async(Vh, a.qkv, row0, ld, 2 * D + col, L, dk, lane);
    load_slot_async(Gh, a.d_ctx, row0, D, col, L, dk, lane);
    // post-dropout context of this lane's (row, 8 columns) pieces, straight from global while the
    // async copies fly: delta_i = sum_d d_ctx * ctx (both carry the same dropout factor)
    float2 ov[4][4];
    float lse[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = ly + 8 * i;
        lse[i] = r < L ? a.lse[(row0 + r) * a.n_heads + h] : 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int d = 8 * lx + 2 * c;
            ov[i][c] = (r < L && d < dk) ? __ldg(reinterpret_cast<const float2*>(a.ctx + (row0 + r) * D + col + d))
                                         : make_float2(0.f, 0.f);
        }
    }
    cp_async_wait_all();
    __syncwarp();

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
        // dO = d_ctx * keep/(1-p), in place: 16 lanes per row, two rows per pass
        __syncwarp();
        const int half = dk >> 1;
        for (int l0 = 0; l0 < L; l0 += 2) {
            const int l = l0 + (lane >> 4), pp = lane & 15;
            if (l < L && pp < half) {
                const int d = pp << 1;
                const int c = col + d;
                const uint32_t keep = (uint32_t)a.cmask[(row0 + l) * a.mask_bytes + (c >> 3)] >> (c & 7);
                float2* p = reinterpret_cast<float2*>(Gh + l * kRowStride + d);
                float2 g = *p;
                g.x = (keep & 1u) ? g.x * a.drop.scale : 0.f;
                g.y = (keep & 2u) ? g.y * a.drop.scale : 0.f;
                *p = g;
            }
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
    const int rows[4] = {ly, ly + 8, ly + 16, ly + 24};
    const ig::Img& im = a.d_qkv_img;
    float* sums = a.d_bias_part + seq * ld;
    float acc[4][8];
    zero_tile(acc);
    tile_atb(acc, Vh, Gh, ly, lx);       // dV[key][d] = sum_row P[row][key] dO[row][d]
    __syncwarp();                        // all reads of P and dO are done
    store_tile(Vh, acc, krows, lx, one); // dV over P
    store_score_tile(Gh, ds, ly, lx);    // dS over dO
    __syncwarp();
    warp_write_slot<true>(Vh, L, dk, row0, 2 * D + col, a.d_qkv, ld, im, nullptr, 0, 1.f, sums, lane);
    zero_tile(acc);
    tile_atb(acc, Gh, Qh, ly, lx);       // dK[key][d] = sum_row dS[row][key] Q[row][d]
    __syncwarp();                        // all reads of Q are done
    store_tile(Qh, acc, krows, lx, one); // dK over Q
    __syncwarp();
    warp_write_slot<true>(Qh, L, dk, row0, D + col, a.d_qkv, ld, im, nullptr, 0, 1.f, sums, lane);
    zero_tile(acc);
    tile_ab(acc, Gh, Kh, ly, lx);        // dQ[row][d] = sum_key dS[row][key] K[key][d]
    __syncwarp();                        // all reads of K are done
    store_tile(Kh, acc, rows, lx, one);  // dQ over K
    __syncwarp();
    warp_write_slot<true>(Kh, L, dk, row0, col, a.d_qkv, ld, im, nullptr, 0, 1.f, sums, lane);
    if (im.hi != nullptr)
        pad_image(im, row0, L, 3 * D, ceil_div(3 * D, 16) * 16, h == a.n_heads - 1, item == n_items - 1, a.M, lane);
}

}  // namespace nrms
