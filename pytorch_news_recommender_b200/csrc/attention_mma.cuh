// attention_mma.cuh — self-attention for sequences of up to 32 tokens (the title encoder) with the
// per-head 32x32x32 products on the tensor cores, forward and backward.  Same math and I/O
// contract as attention.cuh (reference nrms_v0.py:13-23, 46-76, 171-173).
//
// Why: with CUDA-core FMAs a head's products are bound by shared-memory bandwidth (each lane
// re-reads its operands for every k: 3 x 16-byte reads per 32 FMAs, 4 wavefronts each).  With
// warp-level mma.sync.m16n8k16 every operand element is read from shared memory ONCE per product
// into a register fragment, and the math runs on the tensor pipe.  These are tiny batched
// products (30x30x30 per head, 35,200 heads per step), not GEMM tiles: tcgen05's 128-row tiles
// and TMEM round trips do not fit them, the warp-level instruction does.
//
// Precision: the same 3-term bf16 split as the projections (gemm_img.cuh): x = hi + lo,
// A*B ~= Ahi*Bhi + Alo*Bhi + Ahi*Blo, fp32 accumulate — fp32-grade.
//
// One WARP owns one (sequence, head) and never synchronises with another warp.  Operands arrive
// by cp.async into warp-private 32x36 fp32 slots (rows >= L, columns >= d_k zero).  The score
// tile stays in the accumulator fragment layout (row g / g+8, columns 2t, 2t+1 of each 8-wide
// tile; g = lane/4, t = lane%4): the row softmax is two shuffles over a quad, and P / dS feed the
// next product as A fragments straight from registers (the accumulator layout of m16n8 IS the A
// layout of m16n8k16).  P^T / dS^T for the backward's transposed products go through a dead slot.
#pragma once
#include "attention.cuh"

namespace nrms {

constexpr int kTile = 32;               // rows / columns of a head's score tile

// (x, y) -> packed bf16 pairs: hi = {bf16(x) low, bf16(y) high}, lo = the same of the residuals
__device__ __forceinline__ void split_pair(float x, float y, uint32_t& hi, uint32_t& lo) {
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(y), "f"(x));
    const float hx = __uint_as_float(hi << 16), hy = __uint_as_float(hi & 0xffff0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(y - hy), "f"(x - hx));
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

// Warp-level write-out of a finished slot (rows [0,L), columns [0,dk)) to global column gcol0 + c
// of an fp32 matrix (may be null) and/or a split-bf16 image; optional dropout keep bits
// smask[row*8 + (col>>3) - g0]; SUMS adds the per-sequence column sums (bias partials).
// 8 lanes per row (one 16-byte / 4-column read each), four rows per pass of a ROLLED loop: the
// code stays small (the kernels are instruction-fetch sensitive) and every store is coalesced.
// gcol0 and dk are even, so a lane's four columns are two aligned 2-column pairs.
template <bool SUMS, int RS = kRowStride>
__device__ __forceinline__ void warp_write_slot(const float* slot, int L, int dk, long long row0, int gcol0, float* out,
                                                int ld, const ig::Img& img, const uint8_t* smask, int g0,
                                                float drop_scale, float* sums, int lane, int l0 = 0) {
    const int c = (lane & 7) << 2, rsub = lane >> 3;
    const bool act0 = c < dk, act1 = c + 2 < dk;
    const bool has_img = img.hi != nullptr;
    const int col0 = gcol0 + c, col1 = col0 + 2;
    const int ga = col0 >> 3, gb = col1 >> 3;
    const long long ch0 = has_img ? (long long)(ga >> 3) * img.chunk_stride + (col0 & 7) * 2 : 0;
    const long long ch1 = has_img ? (long long)(gb >> 3) * img.chunk_stride + (col1 & 7) * 2 : 0;
#pragma unroll 1
    for (int l = l0 + rsub; l < L; l += 4) {   // rows [l0, L)
        if (!act0) continue;
        float4 v = *reinterpret_cast<const float4*>(slot + l * RS + c);
        if (smask) {
            const uint32_t k0 = (uint32_t)smask[l * 8 + ga - g0] >> (col0 & 7);
            const uint32_t k1 = (uint32_t)smask[l * 8 + gb - g0] >> (col1 & 7);
            v.x = (k0 & 1u) ? v.x * drop_scale : 0.f;
            v.y = (k0 & 2u) ? v.y * drop_scale : 0.f;
            v.z = (k1 & 1u) ? v.z * drop_scale : 0.f;
            v.w = (k1 & 2u) ? v.w * drop_scale : 0.f;
        }
        const long long r = row0 + l;
        if (out) {
            *reinterpret_cast<float2*>(out + r * ld + col0) = make_float2(v.x, v.y);
            if (act1) *reinterpret_cast<float2*>(out + r * ld + col1) = make_float2(v.z, v.w);
        }
        if (has_img) {
            const int r7 = (int)(r & 7);
            const long long rbase = (r >> 3) * 1024 + r7 * 128;
            uint32_t hi, lo;
            split_pair(v.x, v.y, hi, lo);
            long long off = ch0 + rbase + (((ga & 7) ^ r7) << 4);
            *reinterpret_cast<uint32_t*>(img.hi + off) = hi;
            if (img.lo) *reinterpret_cast<uint32_t*>(img.lo + off) = lo;
            if (act1) {
                split_pair(v.z, v.w, hi, lo);
                off = ch1 + rbase + (((gb & 7) ^ r7) << 4);
                *reinterpret_cast<uint32_t*>(img.hi + off) = hi;
                if (img.lo) *reinterpret_cast<uint32_t*>(img.lo + off) = lo;
            }
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
            if (img.lo) *reinterpret_cast<uint32_t*>(img.lo + off) = 0u;
        }
    }
    if (last_item) {
        const int groups = img.chunks * 8;
        const long long npad = img.rows_pad - M;
        for (long long i = lane; i < npad * groups; i += 32) ig::img_store8_zero(img, M + i / groups, (int)(i % groups));
    }
}

__device__ __forceinline__ float quad_max(float v) {
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
    return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

constexpr int kMS = 36;                 // slot row stride (floats): 3 CTAs of 4 warps fit an SM in the backward
constexpr int kMSlot = kTile * kMS;     // floats per slot
constexpr int kMmaWarps = 4;            // backward: 4 slots per warp, 3 CTAs per SM
constexpr int kMmaWarpsFwd = 6;         // forward: 2 slots per warp (Q goes global -> registers), 3 CTAs per SM

__host__ __device__ inline size_t attn_mma_fwd_smem_bytes() {
    return (size_t)kMmaWarpsFwd * (2 * kMSlot * sizeof(float) + kTile * 8);
}
__host__ __device__ inline size_t attn_mma_bwd_smem_bytes() {
    return (size_t)kMmaWarps * 4 * kMSlot * sizeof(float);
}

__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm(   // a pure register operation: not volatile, so the compiler may interleave independent products
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// c += A*B: TERMS = 3 with the 3-term split (fp32-grade), TERMS = 1 plain bf16 (gemm_mode 2)
template <int TERMS>
__device__ __forceinline__ void mma3(float (&c)[4], const uint32_t (&ahi)[4], const uint32_t (&alo)[4],
                                     const uint32_t (&bhi)[2], const uint32_t (&blo)[2]) {
    if (TERMS == 3) {
        mma_bf16(c, alo, bhi[0], bhi[1]);
        mma_bf16(c, ahi, blo[0], blo[1]);
    }
    mma_bf16(c, ahi, bhi[0], bhi[1]);
}

// A fragments (hi/lo) of the 16x16 block (rows 16mt.., k 16ks..) of a row-major [row][k] slot
__device__ __forceinline__ void load_a_frag(uint32_t (&hi)[4], uint32_t (&lo)[4], const float* A, int mt, int ks,
                                            int g, int t) {
    const float* p = A + (16 * mt + g) * kMS + 16 * ks + 2 * t;
    const float2 v0 = *reinterpret_cast<const float2*>(p);
    const float2 v1 = *reinterpret_cast<const float2*>(p + 8 * kMS);
    const float2 v2 = *reinterpret_cast<const float2*>(p + 8);
    const float2 v3 = *reinterpret_cast<const float2*>(p + 8 * kMS + 8);
    split_pair(v0.x, v0.y, hi[0], lo[0]);
    split_pair(v1.x, v1.y, hi[1], lo[1]);
    split_pair(v2.x, v2.y, hi[2], lo[2]);
    split_pair(v3.x, v3.y, hi[3], lo[3]);
}
// B fragments of the 16(k) x 8(n) block from a slot holding B as [n][k] (k contiguous)
__device__ __forceinline__ void load_b_frag_nk(uint32_t (&hi)[2], uint32_t (&lo)[2], const float* B, int nt, int ks,
                                               int g, int t) {
    const float* p = B + (8 * nt + g) * kMS + 16 * ks + 2 * t;
    const float2 v0 = *reinterpret_cast<const float2*>(p);
    const float2 v1 = *reinterpret_cast<const float2*>(p + 8);
    split_pair(v0.x, v0.y, hi[0], lo[0]);
    split_pair(v1.x, v1.y, hi[1], lo[1]);
}
// B fragments of the 16(k) x 8(n) block from a slot holding B as [k][n] (n contiguous)
__device__ __forceinline__ void load_b_frag_kn(uint32_t (&hi)[2], uint32_t (&lo)[2], const float* B, int nt, int ks,
                                               int g, int t) {
    const float* p = B + (16 * ks + 2 * t) * kMS + 8 * nt + g;
    split_pair(p[0], p[kMS], hi[0], lo[0]);
    split_pair(p[8 * kMS], p[9 * kMS], hi[1], lo[1]);
}

// c[mt][nt] += A[32 x 32k] * B^T, both slots row-major over k:  S = Q K^T,  dP = dO V^T
template <int TERMS>
__device__ __forceinline__ void mma_abt(float (&c)[2][4][4], const float* A, const float* B, int g, int t) {
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
        uint32_t ahi[2][4], alo[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) load_a_frag(ahi[mt], alo[mt], A, mt, ks, g, t);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            uint32_t bhi[2], blo[2];
            load_b_frag_nk(bhi, blo, B, nt, ks, g, t);
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) mma3<TERMS>(c[mt][nt], ahi[mt], alo[mt], bhi, blo);
        }
    }
}
// c[mt][nt] += A * B^T with A's fragments already in registers as raw fp32 pairs
// (qa[ks][mt][0..3] = rows g / g+8, k 2t.. / 2t+8.. of the 16x16 block) and B a [n][k] slot:
// S = Q K^T with Q read straight from global memory (no shared-memory slot for Q)
template <int TERMS>
__device__ __forceinline__ void mma_rawA_bt(float (&c)[2][4][4], const float2 (&qa)[2][2][4], const float* B, int g, int t) {
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
        uint32_t ahi[2][4], alo[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int i = 0; i < 4; ++i) split_pair(qa[ks][mt][i].x, qa[ks][mt][i].y, ahi[mt][i], alo[mt][i]);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            uint32_t bhi[2], blo[2];
            load_b_frag_nk(bhi, blo, B, nt, ks, g, t);
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) mma3<TERMS>(c[mt][nt], ahi[mt], alo[mt], bhi, blo);
        }
    }
}
// c[mt][nt] += P * B with P in the accumulator layout (registers) and B a [k][n] slot:
// O = P V,  dQ = dS K
template <int TERMS>
__device__ __forceinline__ void mma_regA_b(float (&c)[2][4][4], const float (&p)[2][4][4], const float* B, int g, int t) {
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
        uint32_t ahi[2][4], alo[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            split_pair(p[mt][2 * ks][0], p[mt][2 * ks][1], ahi[mt][0], alo[mt][0]);          // row g,   k 2t..
            split_pair(p[mt][2 * ks][2], p[mt][2 * ks][3], ahi[mt][1], alo[mt][1]);          // row g+8
            split_pair(p[mt][2 * ks + 1][0], p[mt][2 * ks + 1][1], ahi[mt][2], alo[mt][2]);  // row g,   k 2t+8..
            split_pair(p[mt][2 * ks + 1][2], p[mt][2 * ks + 1][3], ahi[mt][3], alo[mt][3]);  // row g+8
        }
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            uint32_t bhi[2], blo[2];
            load_b_frag_kn(bhi, blo, B, nt, ks, g, t);
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) mma3<TERMS>(c[mt][nt], ahi[mt], alo[mt], bhi, blo);
        }
    }
}
// c[mt][nt] += AT * B with AT a slot holding A^T as [m][k] and B a [k][n] slot:
// dV = P^T dO (AT = P^T[key][row]),  dK = dS^T Q
template <int TERMS>
__device__ __forceinline__ void mma_a_b(float (&c)[2][4][4], const float* AT, const float* B, int g, int t) {
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
        uint32_t ahi[2][4], alo[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) load_a_frag(ahi[mt], alo[mt], AT, mt, ks, g, t);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            uint32_t bhi[2], blo[2];
            load_b_frag_kn(bhi, blo, B, nt, ks, g, t);
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) mma3<TERMS>(c[mt][nt], ahi[mt], alo[mt], bhi, blo);
        }
    }
}
__device__ __forceinline__ void zero_frag(float (&c)[2][4][4]) {
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i) c[mt][nt][i] = 0.f;
}
// accumulator layout -> slot[row][col] (row-major), rows scaled by mul[mt][half]
__device__ __forceinline__ void store_frag(float* slot, const float (&c)[2][4][4], const float (&mul)[2][2], int g, int t) {
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            float* p = slot + (16 * mt + g) * kMS + 8 * nt + 2 * t;
            *reinterpret_cast<float2*>(p) = make_float2(c[mt][nt][0] * mul[mt][0], c[mt][nt][1] * mul[mt][0]);
            *reinterpret_cast<float2*>(p + 8 * kMS) = make_float2(c[mt][nt][2] * mul[mt][1], c[mt][nt][3] * mul[mt][1]);
        }
}
// accumulator layout -> slot[col][row] (transposed)
__device__ __forceinline__ void store_frag_t(float* slot, const float (&c)[2][4][4], int g, int t) {
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            float* p = slot + (8 * nt + 2 * t) * kMS + 16 * mt + g;
            p[0] = c[mt][nt][0];
            p[kMS] = c[mt][nt][1];
            p[8] = c[mt][nt][2];
            p[kMS + 8] = c[mt][nt][3];
        }
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
template <int TERMS>
__global__ void __launch_bounds__(kMmaWarpsFwd * 32) attn_mma_fwd_kernel(const AttnArgs a, long long n_items) {
    extern __shared__ __align__(16) float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long item = (long long)blockIdx.x * kMmaWarpsFwd + warp;
    if (item >= n_items) return;
    const int L = a.L, D = a.D, dk = a.dk;
    const long long seq = item / a.n_heads;
    const int h = (int)(item - seq * a.n_heads);
    float* Kh = smem + (size_t)warp * 2 * kMSlot;   // K, later the output staging
    float* Vh = Kh + kMSlot;
    uint8_t* smask = reinterpret_cast<uint8_t*>(smem + (size_t)kMmaWarpsFwd * 2 * kMSlot) + warp * kTile * 8;
    const long long row0 = seq * L;
    const int ld = 3 * D, col = h * dk;
    const int g = lane >> 2, t = lane & 3;

    load_slot_async<kMS>(Kh, a.qkv, row0, ld, D + col, L, dk, lane);
    load_slot_async<kMS>(Vh, a.qkv, row0, ld, 2 * D + col, L, dk, lane);
    // Q is only ever an A operand: its fragments come straight from global memory into registers
    // (8 rows x 32 contiguous bytes per load instruction), one shared-memory slot less per warp
    float2 qa[2][2][4];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = 16 * mt + 8 * (i & 1) + g, d = 16 * ks + 8 * (i >> 1) + 2 * t;
                qa[ks][mt][i] = (r < L && d < dk) ? __ldg(reinterpret_cast<const float2*>(a.qkv + (row0 + r) * ld + col + d))
                                                  : make_float2(0.f, 0.f);
            }
    const int g0 = col >> 3;
    const bool drop = a.drop.enabled();
    if (drop) {
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

    float s[2][4][4];
    zero_frag(s);
    mma_rawA_bt<TERMS>(s, qa, Kh, g, t);
    // softmax over the keys: a row lives in the 4 lanes of a quad
    float inv[2][2];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
            float m = -INFINITY;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    float& x = s[mt][nt][2 * hf + e];
                    x = (8 * nt + 2 * t + e < L) ? x * a.scale : -INFINITY;   // scores / sqrt(d_k); no key >= L
                    m = fmaxf(m, x);
                }
            m = quad_max(m);
            float sum = 0.f;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    float& x = s[mt][nt][2 * hf + e];
                    x = __expf(x - m);
                    sum += x;
                }
            sum = quad_sum(sum);
            inv[mt][hf] = 1.f / sum;
            const int r = 16 * mt + 8 * hf + g;
            if (t == 0 && r < L) a.lse[(row0 + r) * a.n_heads + h] = m + __logf(sum);
        }
    float o[2][4][4];
    zero_frag(o);
    mma_regA_b<TERMS>(o, s, Vh, g, t);          // unnormalised P straight from registers
    __syncwarp();                        // all reads of K (scores) are long done; V reads done
    store_frag(Kh, o, inv, g, t);        // O = P V / rowsum over the dead K slot
    __syncwarp();
    warp_write_slot<false, kMS>(Kh, L, dk, row0, col, a.ctx, D, a.ctx_img, drop ? smask : nullptr, g0, a.drop.scale,
                                nullptr, lane);
    if (a.ctx_img.hi != nullptr)
        pad_image(a.ctx_img, row0, L, D, a.ctx_img.chunks * 64, h == a.n_heads - 1, item == n_items - 1, a.M, lane, true);
}

// ------------------------------------------------------------------------------------------------
// backward
//   P = exp(scale*Q K^T - lse) ; dP = dO V^T ; dS = scale * P o (dP - delta) ; delta = rowsum(dO o O)
//   dV = P^T dO ; dK = dS^T Q ; dQ = dS K
// ------------------------------------------------------------------------------------------------
template <int TERMS>
__global__ void __launch_bounds__(kMmaWarps * 32) attn_mma_bwd_kernel(const AttnArgs a, long long n_items) {
    extern __shared__ __align__(16) float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long item = (long long)blockIdx.x * kMmaWarps + warp;
    if (item >= n_items) return;
    const int L = a.L, D = a.D, dk = a.dk;
    const long long seq = item / a.n_heads;
    const int h = (int)(item - seq * a.n_heads);
    float* Qh = smem + (size_t)warp * 4 * kMSlot;   // Q  -> dK staging
    float* Kh = Qh + kMSlot;                        // K  -> dQ staging
    float* Vh = Kh + kMSlot;                        // V  -> P^T -> dV staging
    float* Gh = Vh + kMSlot;                        // dO -> dS^T
    const long long row0 = seq * L;
    const int ld = 3 * D, col = h * dk;
    const int g = lane >> 2, t = lane & 3;
    const bool drop = a.drop.enabled() && a.cmask != nullptr;

    load_slot_async<kMS>(Qh, a.qkv, row0, ld, col, L, dk, lane);
    load_slot_async<kMS>(Kh, a.qkv, row0, ld, D + col, L, dk, lane);
    load_slot_async<kMS>(Vh, a.qkv, row0, ld, 2 * D + col, L, dk, lane);
    load_slot_async<kMS>(Gh, a.d_ctx, row0, D, col, L, dk, lane);
    // rows of this lane in the accumulator layout: r(mt,hf) = 16mt + 8hf + g
    float lse[2][2];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
            const int r = 16 * mt + 8 * hf + g;
            lse[mt][hf] = r < L ? a.lse[(row0 + r) * a.n_heads + h] : 0.f;
        }
    cp_async_wait_all();
    __syncwarp();

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
                float2* p = reinterpret_cast<float2*>(Gh + l * kMS + d);
                float2 gv = *p;
                gv.x = (keep & 1u) ? gv.x * a.drop.scale : 0.f;
                gv.y = (keep & 2u) ? gv.y * a.drop.scale : 0.f;
                *p = gv;
            }
        }
        __syncwarp();
    }
    float p[2][4][4], ds[2][4][4];
    zero_frag(p);
    zero_frag(ds);
    mma_abt<TERMS>(p, Qh, Kh, g, t);            // S
    mma_abt<TERMS>(ds, Gh, Vh, g, t);           // dP
    // P, then delta_i = sum_d dO_id O_id = sum_j P_ij dP_ij (O = P V): a row sum over the quad, no
    // extra operand; then dS = scale * P o (dP - delta)
    float delta[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int hf = i >> 1;
                const bool ok = (16 * mt + 8 * hf + g < L) && (8 * nt + 2 * t + (i & 1) < L);
                const float pv = ok ? __expf(p[mt][nt][i] * a.scale - lse[mt][hf]) : 0.f;
                p[mt][nt][i] = pv;
                delta[mt][hf] = fmaf(pv, ds[mt][nt][i], delta[mt][hf]);
            }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) delta[mt][hf] = quad_sum(delta[mt][hf]);
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i) ds[mt][nt][i] = p[mt][nt][i] * (ds[mt][nt][i] - delta[mt][i >> 1]) * a.scale;
    __syncwarp();                        // all reads of V (dP) are done
    store_frag_t(Vh, p, g, t);           // P^T[key][row] over V
    __syncwarp();
    const float one[2][2] = {{1.f, 1.f}, {1.f, 1.f}};
    const ig::Img& im = a.d_qkv_img;
    float* sums = a.d_bias_part ? a.d_bias_part + seq * ld : nullptr;   // null: the weight-gradient GEMM makes them
    float acc[2][4][4];
    zero_frag(acc);
    mma_a_b<TERMS>(acc, Vh, Gh, g, t);          // dV[key][d] = sum_row P^T[key][row] dO[row][d]
    __syncwarp();                        // all reads of P^T and dO are done
    store_frag(Vh, acc, one, g, t);      // dV over P^T
    store_frag_t(Gh, ds, g, t);          // dS^T[key][row] over dO
    __syncwarp();
    warp_write_slot<true, kMS>(Vh, L, dk, row0, 2 * D + col, a.d_qkv, ld, im, nullptr, 0, 1.f, sums, lane);
    zero_frag(acc);
    mma_a_b<TERMS>(acc, Gh, Qh, g, t);          // dK[key][d] = sum_row dS^T[key][row] Q[row][d]
    __syncwarp();                        // all reads of Q are done
    store_frag(Qh, acc, one, g, t);      // dK over Q
    __syncwarp();
    warp_write_slot<true, kMS>(Qh, L, dk, row0, D + col, a.d_qkv, ld, im, nullptr, 0, 1.f, sums, lane);
    zero_frag(acc);
    mma_regA_b<TERMS>(acc, ds, Kh, g, t);       // dQ[row][d] = sum_key dS[row][key] K[key][d], dS from registers
    __syncwarp();                        // all reads of K are done
    store_frag(Kh, acc, one, g, t);      // dQ over K
    __syncwarp();
    warp_write_slot<true, kMS>(Kh, L, dk, row0, col, a.d_qkv, ld, im, nullptr, 0, 1.f, sums, lane);
    if (im.hi != nullptr)
        pad_image(im, row0, L, 3 * D, ceil_div(3 * D, 16) * 16, h == a.n_heads - 1, item == n_items - 1, a.M, lane);
}

}  // namespace nrms
