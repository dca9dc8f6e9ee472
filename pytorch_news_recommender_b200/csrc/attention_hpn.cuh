// attention_hpn.cuh — head-padded attention with SEVERAL warps per (sequence, head), for
// sequences of up to 32 tokens (LT = 32: two warps per item), up to 48 tokens (LT = 48: three warps;
// cfg5's 48-token titles) and up to 64 tokens (LT = 64: four warps; the user encoder's
// 50-click history).  Same math, inputs and outputs as
// attention_hp.cuh (reference nrms_v0.py:13-23, 46-76, 171-173).
//
// Why several warps: shared memory fixes how many items an SM holds (LT = 32 backward: four operand
// pairs = 20 KB per item -> ten items), so with one warp per item an SM runs ten warps and the
// kernel is bound by the latency of each warp's serial instruction stream.  Here warp w of an item
// owns query rows [16w, 16w+16) for S, dP, P, dS, O and dQ, and KEY rows [16w, 16w+16) for dV and
// dK, whose A operands P^T / dS^T range over ALL query rows and are read (ldmatrix.trans) from
// shared P / dS planes that the warps fill together — no product needs a cross-warp reduction.
// The warps of an item meet at a named barrier wherever one reads what another wrote.
//
// Shared memory of an item (LT rows; a "pair" = hi plane + lo plane of LT rows x 80 bytes, which is
// also one LT x 40 fp32 staging tile):
//   forward : K, V pairs (+ keep bits); the output tile is staged over K
//   backward: Q, K, V, dO pairs; dO arrives as an fp32 tile in its pair and is split in place;
//             P goes over V and dS over dO when LT = 32 (a 32 x 32 tile of pairs fits a pair);
//             LT = 64 has separate P / dS planes (rows of LT*2 + 16 bytes); dV, dK, dQ are staged
//             over V, Q, K.
#pragma once
#include "attention_hp.cuh"

namespace nrms {

template <int LT>
struct HpN {
    static_assert(LT == 32 || LT == 48 || LT == 64, "tiles of 32, 48 or 64 rows");
    // rows of a head block in HBM: the 48-row tile (cfg5's 48-token titles) reads the first 48
    // rows of the 64-row blocks that the projection writes for sequences of 33..64 tokens
    static constexpr int BR = LT == 48 ? 64 : LT;
    static constexpr int NW = LT / 16;                 // warps per item
    static constexpr int KSL = LT / 16;                // 16-wide k-steps over query rows / keys
    static constexpr int NTL = LT / 8;                 // 8-wide n-tiles over keys
    static constexpr int PLANE = LT * kHpRowB;         // one bf16 plane of an operand
    static constexpr int PAIR = 2 * PLANE;             // hi + lo == LT x 40 fp32
    static constexpr int PROWB = LT * 2 + 16;          // row bytes of the P / dS planes
    static constexpr int PPLANE = LT * PROWB;
    static constexpr bool SEP = LT > 32;               // P / dS in their own buffers
    static constexpr int ITEM_BWD = 4 * PAIR + (SEP ? 4 * PPLANE : 0);
    // plain-bf16 products (TERMS == 1) never touch the lo planes of P / dS: half the separate planes, which
    // lets a third 64-row item fit an SM (59 KB instead of 77 KB per item)
    static constexpr int ITEM_BWD_HI = 4 * PAIR + (SEP ? 2 * PPLANE : 0);
    __host__ __device__ static constexpr int item_bwd(int terms) { return terms == 3 ? ITEM_BWD : ITEM_BWD_HI; }
    static constexpr int ITEM_FWD = 2 * PAIR + LT * 8;
    static constexpr int ITEMS_BWD = LT == 32 ? 4 : 1; // items per CTA (4 x 2 warps = 256 threads: 128 registers each, no spills; 5 measured 7 % slower)
    static constexpr int ITEMS_FWD = LT == 32 ? 5 : 2;
};

__device__ __forceinline__ void item_bar(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// ---- fragments ---------------------------------------------------------------------------------
// A fragments of rows [m0, m0+16) x KS k-steps.  AT = false: base holds A as [m][k];
// AT = true: base holds A^T as [k][m] (P[row][key] read as P^T, dS likewise).
template <int TERMS, bool AT, int KS>
__device__ __forceinline__ void hpn_load_a(uint32_t (&ah)[KS][4], uint32_t (&al)[KS][4], uint32_t base, int rowB, int lo_off,
                                           int m0, int lane) {
    const int r7 = lane & 7, j0 = (lane >> 3) & 1, j1 = lane >> 4;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
        const uint32_t addr = AT ? base + (16 * ks + r7 + 8 * j1) * rowB + (m0 + 8 * j0) * 2
                                 : base + (m0 + r7 + 8 * j0) * rowB + (16 * ks + 8 * j1) * 2;
        if (AT) ldsm_x4_t(ah[ks], addr); else ldsm_x4(ah[ks], addr);
        if (TERMS == 3) {
            if (AT) ldsm_x4_t(al[ks], addr + lo_off); else ldsm_x4(al[ks], addr + lo_off);
        }
    }
}
// accumulator tile 16 x (16 KS) -> A fragments (hi / lo)
template <int TERMS, int KS>
__device__ __forceinline__ void hpn_split_acc(uint32_t (&ah)[KS][4], uint32_t (&al)[KS][4], const float (&p)[2 * KS][4]) {
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
        split_pair(p[2 * ks][0], p[2 * ks][1], ah[ks][0], al[ks][0]);
        split_pair(p[2 * ks][2], p[2 * ks][3], ah[ks][1], al[ks][1]);
        split_pair(p[2 * ks + 1][0], p[2 * ks + 1][1], ah[ks][2], al[ks][2]);
        split_pair(p[2 * ks + 1][2], p[2 * ks + 1][3], ah[ks][3], al[ks][3]);
    }
}
// the same fragments -> rows [m0, m0+16) of row-major planes
template <int TERMS, int KS>
__device__ __forceinline__ void hpn_store_a(uint32_t base, int rowB, int lo_off, const uint32_t (&ah)[KS][4],
                                            const uint32_t (&al)[KS][4], int m0, int g, int t) {
#pragma unroll
    for (int ks = 0; ks < KS; ++ks)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t addr = base + (m0 + g + 8 * (i & 1)) * rowB + (16 * ks + 8 * (i >> 1) + 2 * t) * 2;
            sts32(addr, ah[ks][i]);
            if (TERMS == 3) sts32(addr + lo_off, al[ks][i]);
        }
}
template <int TERMS>
__device__ __forceinline__ void hpn_mma3(float (&c)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4], uint32_t bh0, uint32_t bh1,
                                         uint32_t bl0, uint32_t bl1) {
    if (TERMS == 3) {
        mma_bf16(c, al, bh0, bh1);
        mma_bf16(c, ah, bl0, bl1);
    }
    mma_bf16(c, ah, bh0, bh1);
}
// c[nt] += A[16 x 32] * B^T, B an operand pair holding [n][k] (k = the 32 head columns): S, dP
template <int TERMS, int NT>
__device__ __forceinline__ void hpn_mma_nk(float (&c)[NT][4], const uint32_t (&ah)[2][4], const uint32_t (&al)[2][4], uint32_t pair,
                                           int lo_off, int lane) {
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        const uint32_t addr = pair + (8 * nt + (lane & 7)) * kHpRowB + (lane >> 3) * 16;
        uint32_t bh[4], bl[4] = {0u, 0u, 0u, 0u};
        ldsm_x4(bh, addr);
        if (TERMS == 3) ldsm_x4(bl, addr + lo_off);
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) hpn_mma3<TERMS>(c[nt], ah[ks], al[ks], bh[2 * ks], bh[2 * ks + 1], bl[2 * ks], bl[2 * ks + 1]);
    }
}
// c[nt] += A[16 x 16 KS] * B, B an operand pair holding [k][n] (n = the 32 head columns): O, dV, dK, dQ
template <int TERMS, int KS>
__device__ __forceinline__ void hpn_mma_kn(float (&c)[4][4], const uint32_t (&ah)[KS][4], const uint32_t (&al)[KS][4], uint32_t pair,
                                           int lo_off, int lane) {
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int kh = 0; kh < KS / 2; ++kh) {
            const uint32_t addr = pair + (lane + 32 * kh) * kHpRowB + nt * 16;
            uint32_t bh[4], bl[4] = {0u, 0u, 0u, 0u};
            ldsm_x4_t(bh, addr);
            if (TERMS == 3) ldsm_x4_t(bl, addr + lo_off);
#pragma unroll
            for (int j = 0; j < 2; ++j)
                hpn_mma3<TERMS>(c[nt], ah[2 * kh + j], al[2 * kh + j], bh[2 * j], bh[2 * j + 1], bl[2 * j], bl[2 * j + 1]);
        }
    if (KS & 1) {
        // odd number of 16-wide k-steps (48-row tiles): the last one is a 16-row ldmatrix.x2
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            const uint32_t addr = pair + (16 * (KS - 1) + (lane & 15)) * kHpRowB + nt * 16;
            uint32_t bh[2], bl[2] = {0u, 0u};
            ldsm_x2_t(bh, addr);
            if (TERMS == 3) ldsm_x2_t(bl, addr + lo_off);
            hpn_mma3<TERMS>(c[nt], ah[KS - 1], al[KS - 1], bh[0], bh[1], bl[0], bl[1]);
        }
    }
}
template <int N>
__device__ __forceinline__ void hpn_zero(float (&c)[N][4]) {
#pragma unroll
    for (int nt = 0; nt < N; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) c[nt][i] = 0.f;
}
// accumulator rows [m0, m0+16) x 32 columns -> the fp32 staging tile, rows scaled by mul[half]
__device__ __forceinline__ void hpn_stage(float* tile, const float (&c)[4][4], float mul0, float mul1, int m0, int g, int t) {
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
        float* p = tile + (m0 + g) * kHpStage + 8 * nt + 2 * t;
        *reinterpret_cast<float2*>(p) = make_float2(c[nt][0] * mul0, c[nt][1] * mul0);
        *reinterpret_cast<float2*>(p + 8 * kHpStage) = make_float2(c[nt][2] * mul1, c[nt][3] * mul1);
    }
}
// rows [m0, m0+16) of the staging tile -> image columns [gcol0, gcol0 + 32); a lane owns one
// 16-byte unit of a row per pass
__device__ __forceinline__ void hpn_write_img(const float* tile, int m0, int L, long long row0, int gcol0, const ig::Img& img,
                                              int lane) {
    const int r8 = lane >> 2, u = lane & 3;
    const int gg = (gcol0 >> 3) + u;
    const long long cbase = (long long)(gg >> 3) * img.chunk_stride;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int l = m0 + r8 + 8 * i;
        if (l < L) {
            const float4 v0 = *reinterpret_cast<const float4*>(tile + l * kHpStage + 8 * u);
            const float4 v1 = *reinterpret_cast<const float4*>(tile + l * kHpStage + 8 * u + 4);
            uint32_t hi[4], lo[4];
            split_pair(v0.x, v0.y, hi[0], lo[0]);
            split_pair(v0.z, v0.w, hi[1], lo[1]);
            split_pair(v1.x, v1.y, hi[2], lo[2]);
            split_pair(v1.z, v1.w, hi[3], lo[3]);
            const long long r = row0 + l;
            const int r7 = (int)(r & 7);
            const long long off = cbase + (r >> 3) * 1024 + r7 * 128 + (((gg & 7) ^ r7) << 4);
            *reinterpret_cast<uint4*>(img.hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            if (img.lo) *reinterpret_cast<uint4*>(img.lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
    }
}
// ---- global -> shared --------------------------------------------------------------------------
// rows [r0, r0+16) of one head's operand block -> the pair; rows >= L zero-filled
template <bool LO>
__device__ __forceinline__ void hpn_load_rows(uint32_t pair, int lo_off, const uint16_t* hi, const uint16_t* lo, long long blk,
                                              int r0, int L, int lane) {
    const int r8 = lane >> 2, u = lane & 3;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int r = r0 + r8 + 8 * i;
        const bool ok = r < L;
        const long long off = blk + (ok ? r : 0) * 32 + u * 8;
        const uint32_t dst = pair + r * kHpRowB + u * 16;
        cp_async16_zfill(dst, hi + off, ok ? 16u : 0u);
        if (LO) cp_async16_zfill(dst + lo_off, lo + off, ok ? 16u : 0u);
    }
}
// rows [r0, min(r0+16, L)) x columns [col, col+dk) of d_ctx -> the fp32 tile (8-byte cp.async)
__device__ __forceinline__ void hpn_request_do(uint32_t tile, const float* d_ctx, long long row0, int D, int col, int r0, int L,
                                               int dk, int lane) {
    const int pp = lane & 15, par = lane >> 4;
    const int r1 = r0 + 16 < L ? r0 + 16 : L;
    if (2 * pp < dk) {
        const float* gp = d_ctx + (row0 + r0 + par) * D + col + 2 * pp;
        uint32_t dst = tile + ((r0 + par) * kHpStage + 2 * pp) * 4;
        for (int l = r0 + par; l < r1; l += 2) {
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(gp) : "memory");
            dst += 2 * kHpStage * 4;
            gp += 2 * D;
        }
    }
}
template <bool LO>
__device__ __forceinline__ void hpn_prefetch_blocks(const AttnArgs& a, long long seq, int h, int rows, int lane, int blk_rows = 0) {
    if (lane < (LO ? 6 : 3)) {
        const int plane = lane / 3, which = lane - plane * 3;
        const uint16_t* p = (plane ? a.qkv_lo : a.qkv_hi) + hp_block_off(seq, which, h, a.n_heads, blk_rows ? blk_rows : rows);
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(rows * 64) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
template <int TERMS, int LT>
__global__ void __launch_bounds__(HpN<LT>::ITEMS_FWD* HpN<LT>::NW * 32, LT == 48 ? (TERMS == 1 ? 4 : 3) : 2) attn_hpn_fwd_kernel(const AttnArgs a, long long n_items) {
    using C = HpN<LT>;
    extern __shared__ __align__(16) float smem[];
    uint8_t* sm = reinterpret_cast<uint8_t*>(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slot = warp / C::NW, w = warp % C::NW, m0 = 16 * w, bar = 1 + slot;
    const int g = lane >> 2, t = lane & 3;
    const int L = a.L, D = a.D, dk = a.dk;
    const long long stride = (long long)gridDim.x * C::ITEMS_FWD;
    uint8_t* Kb = sm + (size_t)slot * C::ITEM_FWD;
    const uint32_t Ks = (uint32_t)__cvta_generic_to_shared(Kb), Vs = Ks + C::PAIR;
    uint8_t* smask = Kb + 2 * C::PAIR;
    float* const stage = reinterpret_cast<float*>(Kb);
    const bool drop = a.drop.enabled();
    for (long long item = (long long)blockIdx.x * C::ITEMS_FWD + slot; item < n_items; item += stride) {
        const long long seq = item / a.n_heads;
        const int h = (int)(item - seq * a.n_heads);
        const long long row0 = seq * L;
        const int col = h * dk, g0 = col >> 3;
        hpn_load_rows<TERMS == 3>(Ks, C::PLANE, a.qkv_hi, a.qkv_lo, hp_block_off(seq, 1, h, a.n_heads, C::BR), m0, L, lane);
        hpn_load_rows<TERMS == 3>(Vs, C::PLANE, a.qkv_hi, a.qkv_lo, hp_block_off(seq, 2, h, a.n_heads, C::BR), m0, L, lane);
        // Q is only ever an A operand: the own rows' fragments come straight from the global planes
        uint32_t qh[2][4], ql[2][4];
        const long long qblk = hp_block_off(seq, 0, h, a.n_heads, C::BR);
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = m0 + 8 * (i & 1) + g, d = 16 * ks + 8 * (i >> 1) + 2 * t;
                const long long off = qblk + r * 32 + d;
                qh[ks][i] = r < L ? __ldg(reinterpret_cast<const uint32_t*>(a.qkv_hi + off)) : 0u;
                ql[ks][i] = (TERMS == 3 && r < L) ? __ldg(reinterpret_cast<const uint32_t*>(a.qkv_lo + off)) : 0u;
            }
        const int r_end = m0 + 16 < L ? m0 + 16 : L;
        if (drop) {
            // keep bits of the own rows, only the ng (<= 5) eight-column groups this head touches
            const int ng = ((col + dk + 7) >> 3) - g0;
            for (int it = lane; it < (r_end - m0) * ng; it += 32) {
                const int l = m0 + it / ng, gi = it % ng;
                const uint32_t keep = a.drop.keep8(kDropContext, (uint64_t)(row0 + l), (uint32_t)(g0 + gi));
                smask[l * 8 + gi] = (uint8_t)keep;
                if (a.cmask && g0 + gi < a.mask_bytes) a.cmask[(row0 + l) * a.mask_bytes + g0 + gi] = (uint8_t)keep;
            }
        }
        cp_async_wait_all();
        item_bar(bar, C::NW * 32);                          // K and V complete
        if (w == 0 && item + stride < n_items) {
            const long long nseq = (item + stride) / a.n_heads;
            hpn_prefetch_blocks<TERMS == 3>(a, nseq, (int)(item + stride - nseq * a.n_heads), LT, lane, C::BR);
        }
        float s[C::NTL][4];
        hpn_zero(s);
        hpn_mma_nk<TERMS, C::NTL>(s, qh, ql, Ks, C::PLANE, lane);
        // softmax over the keys: a row lives in the 4 lanes of a quad
        float inv[2];
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
            float m = -INFINITY;
#pragma unroll
            for (int nt = 0; nt < C::NTL; ++nt)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    float& x = s[nt][2 * hf + e];
                    x = (8 * nt + 2 * t + e < L) ? x * a.scale : -INFINITY;   // scores / sqrt(d_k); no key >= L
                    m = fmaxf(m, x);
                }
            m = quad_max(m);
            float sum = 0.f;
#pragma unroll
            for (int nt = 0; nt < C::NTL; ++nt)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    float& x = s[nt][2 * hf + e];
                    x = __expf(x - m);
                    sum += x;
                }
            sum = quad_sum(sum);
            inv[hf] = 1.f / sum;
            const int r = m0 + 8 * hf + g;
            if (t == 0 && r < L) a.lse[(row0 + r) * a.n_heads + h] = m + __logf(sum);
        }
        float o[4][4];
        hpn_zero(o);
        {
            uint32_t ph[C::KSL][4], pl[C::KSL][4];
            hpn_split_acc<TERMS, C::KSL>(ph, pl, s);        // unnormalised P straight from registers
            hpn_mma_kn<TERMS, C::KSL>(o, ph, pl, Vs, C::PLANE, lane);
        }
        item_bar(bar, C::NW * 32);                          // every warp is done reading K: the tile goes over it
        hpn_stage(stage, o, inv[0], inv[1], m0, g, t);      // O = P V / rowsum, own rows
        __syncwarp();
        warp_write_slot<false, kHpStage>(stage, r_end, dk, row0, col, a.ctx, D, a.ctx_img, drop ? smask : nullptr, g0,
                                         a.drop.scale, nullptr, lane, m0);
        if (a.ctx_img.hi != nullptr && w == 0)
            pad_image(a.ctx_img, row0, L, D, a.ctx_img.chunks * 64, h == a.n_heads - 1, item == n_items - 1, a.M, lane, true);
        item_bar(bar, C::NW * 32);                          // staging reads are done before the next item's copies land
    }
}

// ------------------------------------------------------------------------------------------------
// backward
//   P = exp(scale*Q K^T - lse) ; dP = dO V^T ; dS = scale * P o (dP - delta) ; delta = rowsum(P o dP)
//   dV = P^T dO ; dK = dS^T Q ; dQ = dS K        -> d_qkv image, head-padded column order
// ------------------------------------------------------------------------------------------------
template <int TERMS, int LT>
__global__ void __launch_bounds__(HpN<LT>::ITEMS_BWD* HpN<LT>::NW * 32, LT == 48 ? (TERMS == 1 ? 5 : 4) : (LT == 64 && TERMS == 1) ? 3 : 2)
attn_hpn_bwd_kernel(const AttnArgs a, long long n_items) {
    using C = HpN<LT>;
    extern __shared__ __align__(16) float smem[];
    uint8_t* sm = reinterpret_cast<uint8_t*>(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slot = warp / C::NW, w = warp % C::NW, m0 = 16 * w, bar = 1 + slot;
    constexpr int NT = C::NW * 32;
    const int g = lane >> 2, t = lane & 3;
    const int L = a.L, D = a.D, dk = a.dk, DP = 32 * a.n_heads;
    const long long stride = (long long)gridDim.x * C::ITEMS_BWD;
    uint8_t* Qb = sm + (size_t)slot * C::item_bwd(TERMS);
    const uint32_t Qs = (uint32_t)__cvta_generic_to_shared(Qb);   // Q  -> dK staging
    const uint32_t Ks = Qs + C::PAIR;                             // K  -> dQ staging
    const uint32_t Vs = Ks + C::PAIR;                             // V  (-> P when LT = 32) -> dV staging
    const uint32_t Gs = Vs + C::PAIR;                             // dO fp32 tile -> dO planes (-> dS when LT = 32)
    const uint32_t Ps = C::SEP ? Gs + C::PAIR : Vs;               // P[row][key] planes
    const uint32_t Ss = C::SEP ? Ps + (TERMS == 3 ? 2 : 1) * C::PPLANE : Gs;   // dS[row][key] planes
    float* const stageQ = reinterpret_cast<float*>(Qb);
    float* const stageK = reinterpret_cast<float*>(Qb + C::PAIR);
    float* const stageV = reinterpret_cast<float*>(Qb + 2 * C::PAIR);
    const float* const tileG = reinterpret_cast<const float*>(Qb + 3 * C::PAIR);
    const bool drop = a.drop.enabled() && a.cmask != nullptr;
    const ig::Img& im = a.d_qkv_img;
    // dO of the NEXT item is requested (cp.async into the dO pair, dead after the dK product) while the
    // current item still has its dQ product and two write-outs to do, and consumed at the top of the
    // next iteration: its DRAM latency never stalls the warps
    {
        const long long item0 = (long long)blockIdx.x * C::ITEMS_BWD + slot;
        if (item0 < n_items) {
            const long long seq0 = item0 / a.n_heads;
            hpn_request_do(Gs, a.d_ctx, seq0 * L, D, (int)(item0 - seq0 * a.n_heads) * dk, m0, L, dk, lane);
        }
    }
    for (long long item = (long long)blockIdx.x * C::ITEMS_BWD + slot; item < n_items; item += stride) {
        const long long seq = item / a.n_heads;
        const int h = (int)(item - seq * a.n_heads);
        const long long row0 = seq * L;
        const int col = h * dk, colp = h * 32;
        const bool has_next = item + stride < n_items;
        const long long nseq = (item + stride) / a.n_heads;
        const int nh = (int)(item + stride - nseq * a.n_heads);
        hpn_load_rows<TERMS == 3>(Qs, C::PLANE, a.qkv_hi, a.qkv_lo, hp_block_off(seq, 0, h, a.n_heads, C::BR), m0, L, lane);
        hpn_load_rows<TERMS == 3>(Ks, C::PLANE, a.qkv_hi, a.qkv_lo, hp_block_off(seq, 1, h, a.n_heads, C::BR), m0, L, lane);
        hpn_load_rows<TERMS == 3>(Vs, C::PLANE, a.qkv_hi, a.qkv_lo, hp_block_off(seq, 2, h, a.n_heads, C::BR), m0, L, lane);
        float lse[2];
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
            const int r = m0 + 8 * hf + g;
            lse[hf] = r < L ? a.lse[(row0 + r) * a.n_heads + h] : 0.f;
        }
        cp_async_wait_all();
        item_bar(bar, NT);                                  // Q K V and the dO tile are visible to every warp
        // dO (d_ctx carries the context-dropout mask already when the tensor-core data-gradient GEMM made
        // it): own rows of the fp32 tile -> fragments, split once, kept as the A operand of dP and stored
        // over the tile as [row][d] planes for dV's B operand
        uint32_t gh[2][4], gl[2][4];
        {
            float2 v[2][4];
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int r = m0 + 8 * (i & 1) + g, d = 16 * ks + 8 * (i >> 1) + 2 * t;
                    v[ks][i] = (r < L && d < dk) ? *reinterpret_cast<const float2*>(tileG + r * kHpStage + d) : make_float2(0.f, 0.f);
                    if (drop && r < L && d < dk) {
                        const int c = col + d;
                        const uint32_t keep = (uint32_t)__ldg(a.cmask + (row0 + r) * a.mask_bytes + (c >> 3)) >> (c & 7);
                        v[ks][i].x = (keep & 1u) ? v[ks][i].x * a.drop.scale : 0.f;
                        v[ks][i].y = (keep & 2u) ? v[ks][i].y * a.drop.scale : 0.f;
                    }
                }
            item_bar(bar, NT);                              // every warp holds its fp32 dO before the planes overwrite the tile
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
#pragma unroll
                for (int i = 0; i < 4; ++i) split_pair(v[ks][i].x, v[ks][i].y, gh[ks][i], gl[ks][i]);
        }
        hpn_store_a<TERMS, 2>(Gs, kHpRowB, C::PLANE, gh, gl, m0, g, t);
        if (has_next && w == 0) {
            hpn_prefetch_blocks<TERMS == 3>(a, nseq, nh, LT, lane, C::BR);
            if (lane < L) prefetch_l2(a.lse + (nseq * L + lane) * a.n_heads + nh);
        }
        float p[C::NTL][4], ds[C::NTL][4];
        hpn_zero(p);
        hpn_zero(ds);
        {
            uint32_t qh[2][4], ql[2][4];
            hpn_load_a<TERMS, false, 2>(qh, ql, Qs, kHpRowB, C::PLANE, m0, lane);
            hpn_mma_nk<TERMS, C::NTL>(p, qh, ql, Ks, C::PLANE, lane);    // S  (own query rows x all keys)
        }
        hpn_mma_nk<TERMS, C::NTL>(ds, gh, gl, Vs, C::PLANE, lane);       // dP
        float delta[2] = {0.f, 0.f};
#pragma unroll
        for (int nt = 0; nt < C::NTL; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int hf = i >> 1;
                const bool ok = (m0 + 8 * hf + g < L) && (8 * nt + 2 * t + (i & 1) < L);
                const float pv = ok ? __expf(p[nt][i] * a.scale - lse[hf]) : 0.f;
                p[nt][i] = pv;
                delta[hf] = fmaf(pv, ds[nt][i], delta[hf]);
            }
        delta[0] = quad_sum(delta[0]);
        delta[1] = quad_sum(delta[1]);
#pragma unroll
        for (int nt = 0; nt < C::NTL; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i) ds[nt][i] = p[nt][i] * (ds[nt][i] - delta[i >> 1]) * a.scale;
        uint32_t sh[C::KSL][4], sl[C::KSL][4];              // dS fragments: A operand of dQ, stored for dK
        hpn_split_acc<TERMS, C::KSL>(sh, sl, ds);
        {
            uint32_t ph[C::KSL][4], pl[C::KSL][4];
            hpn_split_acc<TERMS, C::KSL>(ph, pl, p);
            if (!C::SEP) item_bar(bar, NT);                 // every warp is done reading V (dP): P goes over it
            hpn_store_a<TERMS, C::KSL>(Ps, C::PROWB, C::PPLANE, ph, pl, m0, g, t);
        }
        item_bar(bar, NT);                                  // P and the dO planes are complete
        float acc[4][4];
        uint32_t ah[C::KSL][4], al[C::KSL][4];
        hpn_zero(acc);
        hpn_load_a<TERMS, true, C::KSL>(ah, al, Ps, C::PROWB, C::PPLANE, m0, lane);   // P^T, own keys x all query rows
        hpn_mma_kn<TERMS, C::KSL>(acc, ah, al, Gs, C::PLANE, lane);                   // dV[own keys][d]
        if (!C::SEP) item_bar(bar, NT);                     // every warp is done reading P and dO: dV / dS go over them
        hpn_stage(stageV, acc, 1.f, 1.f, m0, g, t);         // own rows of dV
        hpn_store_a<TERMS, C::KSL>(Ss, C::PROWB, C::PPLANE, sh, sl, m0, g, t);        // own rows of dS[row][key]
        item_bar(bar, NT);                                  // dS complete (the own staging rows are visible)
        hpn_write_img(stageV, m0, L, row0, 2 * DP + colp, im, lane);
        hpn_zero(acc);
        hpn_load_a<TERMS, true, C::KSL>(ah, al, Ss, C::PROWB, C::PPLANE, m0, lane);   // dS^T, own keys x all query rows
        hpn_mma_kn<TERMS, C::KSL>(acc, ah, al, Qs, C::PLANE, lane);                   // dK[own keys][d]
        item_bar(bar, NT);                                  // every warp is done reading Q, dS and the dO planes
        if (has_next) hpn_request_do(Gs, a.d_ctx, nseq * L, D, nh * dk, m0, L, dk, lane);   // next item's dO, own rows
        hpn_stage(stageQ, acc, 1.f, 1.f, m0, g, t);         // own rows of dK over Q
        __syncwarp();
        hpn_write_img(stageQ, m0, L, row0, DP + colp, im, lane);
        hpn_zero(acc);
        hpn_mma_kn<TERMS, C::KSL>(acc, sh, sl, Ks, C::PLANE, lane);                   // dQ[own rows][d]
        item_bar(bar, NT);                                  // every warp is done reading K
        hpn_stage(stageK, acc, 1.f, 1.f, m0, g, t);         // own rows of dQ over K
        __syncwarp();
        hpn_write_img(stageK, m0, L, row0, colp, im, lane);
        if (w == 0) pad_image(im, row0, L, 0, 0, false, item == n_items - 1, a.M, lane);
        item_bar(bar, NT);                                  // staging reads are done before the next item's copies land
    }
}

}  // namespace nrms
