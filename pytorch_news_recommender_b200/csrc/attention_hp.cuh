// attention_hp.cuh — self-attention for sequences of up to 32 tokens over the HEAD-PADDED
// split-bf16 Q|K|V planes the projection GEMM writes (gemm_img.cuh: EPI_BIAS_SPLIT, hp_unpad).
// Same math and outputs as attention_mma.cuh (reference nrms_v0.py:13-23, 46-76, 171-173).
//
// What changes against attention_mma.cuh is how operands reach the tensor pipe:
//   * Q|K|V arrive as two bf16 planes (hi = bf16(x), lo = bf16(x - hi)), every head 32 columns wide
//     and 64-byte aligned, so a head's operand is 32 rows x 4 sixteen-byte units per plane:
//     cp.async.cg of 16 bytes with zero-fill for rows >= L (8 per lane per operand instead of 14
//     eight-byte copies plus zero-fill stores), and NO fp32 -> bf16 split inside the kernel;
//   * every fragment is one ldmatrix.x4 (4 per operand plane per product) instead of 16-32 scalar
//     shared-memory loads plus ~7 ALU instructions per pair for the split; operands that the
//     backward needs transposed (P^T, dS^T, and V / dO / Q / K as [k][n] B operands) use
//     ldmatrix.trans on the same row-major planes — no transposed scatter stores;
//   * the d_qkv image keeps the padded column order, so a head's output row is 4 aligned 16-byte
//     units per plane: one lane stores a whole unit (8 rows x 64 B per store instruction).
// Measured instruction count per (sequence, head): backward 5,775 -> ~2,100 warp instructions.
//
// Slot geometry: a plane slot is 32 rows x 80 bytes (64 data + 16 pad: the eight row addresses of
// an ldmatrix phase fall in eight different 16-byte bank groups); a PAIR (hi plane, lo plane) is
// 5,120 bytes, which is also one 32 x 40 fp32 staging tile for the write-out.
#pragma once
#include "attention_mma.cuh"

namespace nrms {

constexpr int kHpRowB = 80;
constexpr int kHpPlaneB = kTile * kHpRowB;     // 2,560
constexpr int kHpPairB = 2 * kHpPlaneB;        // 5,120
constexpr int kHpStage = 40;                   // floats per staging row (32 x 40 x 4 = one pair)
constexpr int kHpFwdWarps = 8;                 // forward : K, V pairs per warp        -> 2 CTAs / SM
constexpr int kHpBwdWarps = 5;                 // backward: Q, K, V, dO pairs per warp -> 2 CTAs / SM

__host__ __device__ inline size_t attn_hp_fwd_smem_bytes() {
    return (size_t)kHpFwdWarps * (2 * kHpPairB + kTile * 8);
}
__host__ __device__ inline size_t attn_hp_bwd_smem_bytes() { return (size_t)kHpBwdWarps * 4 * kHpPairB; }

__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(addr)
                 : "memory");
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(addr)
                 : "memory");
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// element offset of the 32 x 32 block of (sequence, which = Q/K/V, head) in a plane: the planes are
// HEAD-BLOCKED, [n_seq][3][h] blocks of 32 rows x 32 bf16 (2 KB contiguous per operand), so that a
// warp's operand is one contiguous 2 KB read instead of 30 64-byte pieces at a 1,920-byte stride
__host__ __device__ __forceinline__ long long hp_block_off(long long seq, int which, int head, int n_heads) {
    return ((seq * 3 + which) * n_heads + head) * 1024;
}
// L2 prefetch of the NEXT item's operand blocks (a persistent warp walks items with a fixed stride):
// one bulk prefetch per 2 KB block, issued while the current item computes, so the next item's
// cp.async / loads hit L2 instead of waiting on DRAM
template <bool LO>
__device__ __forceinline__ void hp_prefetch_blocks(const AttnArgs& a, long long seq, int h, int first_which, int lane) {
    const int nb = 3 - first_which;                     // forward skips nothing (Q K V), same for backward
    if (lane < (LO ? 2 : 1) * nb) {
        const int plane = lane / nb, which = first_which + lane - plane * nb;
        const uint16_t* p = (plane ? a.qkv_lo : a.qkv_hi) + hp_block_off(seq, which, h, a.n_heads);
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(2048) : "memory");
    }
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// one head's 32 x 32 operand block -> a pair; rows >= L zero
template <bool LO>
__device__ __forceinline__ void hp_load_pair(uint32_t pair, const uint16_t* hi, const uint16_t* lo, long long blk,
                                             int L, int lane) {
    const int r8 = lane >> 2, u = lane & 3;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = r8 + 8 * i;
        const bool ok = r < L;
        const long long off = blk + (ok ? r : 0) * 32 + u * 8;
        const uint32_t dst = pair + r * kHpRowB + u * 16;
        cp_async16_zfill(dst, hi + off, ok ? 16u : 0u);
        if (LO) cp_async16_zfill(dst + kHpPlaneB, lo + off, ok ? 16u : 0u);
    }
}

// A fragments of all four 16x16 blocks of a 32 x 32 operand held in a pair.
//   AT = false: the pair holds A as [m][k]   (S = Q K^T: A = Q)
//   AT = true : the pair holds A^T as [k][m] (dV = P^T dO: the pair holds P[row][key]; dK = dS^T Q)
template <int TERMS, bool AT>
__device__ __forceinline__ void hp_load_a(uint32_t (&ah)[2][2][4], uint32_t (&al)[2][2][4], uint32_t pair, int lane) {
    const int r7 = lane & 7, j0 = (lane >> 3) & 1, j1 = lane >> 4;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            const uint32_t addr = AT ? pair + (16 * ks + r7 + 8 * j1) * kHpRowB + (16 * mt + 8 * j0) * 2
                                     : pair + (16 * mt + r7 + 8 * j0) * kHpRowB + (16 * ks + 8 * j1) * 2;
            if (AT) ldsm_x4_t(ah[ks][mt], addr); else ldsm_x4(ah[ks][mt], addr);
            if (TERMS == 3) {
                if (AT) ldsm_x4_t(al[ks][mt], addr + kHpPlaneB); else ldsm_x4(al[ks][mt], addr + kHpPlaneB);
            }
        }
}
// A fragments (hi / lo) from a 32 x 32 tile in the accumulator layout (P, dS): the m16n8
// accumulator layout IS the m16n8k16 A layout, pair by pair
template <int TERMS>
__device__ __forceinline__ void hp_split_acc(uint32_t (&ah)[2][2][4], uint32_t (&al)[2][2][4], const float (&p)[2][4][4]) {
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            split_pair(p[mt][2 * ks][0], p[mt][2 * ks][1], ah[ks][mt][0], al[ks][mt][0]);          // row g,   k 2t..
            split_pair(p[mt][2 * ks][2], p[mt][2 * ks][3], ah[ks][mt][1], al[ks][mt][1]);          // row g+8
            split_pair(p[mt][2 * ks + 1][0], p[mt][2 * ks + 1][1], ah[ks][mt][2], al[ks][mt][2]);  // row g,   k 2t+8..
            split_pair(p[mt][2 * ks + 1][2], p[mt][2 * ks + 1][3], ah[ks][mt][3], al[ks][mt][3]);  // row g+8
        }
}
// the same fragments -> a pair holding the tile row-major [row][col] (register i of block
// (ks, mt) = row 16mt + g + 8(i&1), columns 16ks + 8(i>>1) + 2t, +1)
template <int TERMS>
__device__ __forceinline__ void hp_store_a(uint32_t pair, const uint32_t (&ah)[2][2][4], const uint32_t (&al)[2][2][4],
                                           int g, int t) {
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint32_t addr = pair + (16 * mt + g + 8 * (i & 1)) * kHpRowB + (16 * ks + 8 * (i >> 1) + 2 * t) * 2;
                sts32(addr, ah[ks][mt][i]);
                if (TERMS == 3) sts32(addr + kHpPlaneB, al[ks][mt][i]);
            }
}
// c[mt][nt] += A * B over the whole 32 x 32 x 32 product, A in registers, B in a pair:
//   BT = false: the pair holds B as [n][k]  (S = Q K^T: K;  dP = dO V^T: V)
//   BT = true : the pair holds B as [k][n]  (O = P V: V;  dV: dO;  dK: Q;  dQ: K)
template <int TERMS, bool BT>
__device__ __forceinline__ void hp_mma(float (&c)[2][4][4], const uint32_t (&ah)[2][2][4], const uint32_t (&al)[2][2][4],
                                       uint32_t pair, int lane) {
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
        // one ldmatrix.x4 = the (k 0-7, k 8-15) halves of both k-steps for this 8-wide n block
        const uint32_t addr = BT ? pair + lane * kHpRowB + nt * 16 : pair + (8 * nt + (lane & 7)) * kHpRowB + (lane >> 3) * 16;
        uint32_t bh[4], bl[4];
        if (BT) ldsm_x4_t(bh, addr); else ldsm_x4(bh, addr);
        if (TERMS == 3) {
            if (BT) ldsm_x4_t(bl, addr + kHpPlaneB); else ldsm_x4(bl, addr + kHpPlaneB);
        }
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            if (TERMS == 3) {
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) mma_bf16(c[mt][nt], al[ks][mt], bh[2 * ks], bh[2 * ks + 1]);
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) mma_bf16(c[mt][nt], ah[ks][mt], bl[2 * ks], bl[2 * ks + 1]);
            }
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) mma_bf16(c[mt][nt], ah[ks][mt], bh[2 * ks], bh[2 * ks + 1]);
        }
    }
}
// accumulator layout -> fp32 staging tile (row stride kHpStage), rows scaled by mul[mt][half]
__device__ __forceinline__ void hp_stage(float* tile, const float (&c)[2][4][4], const float (&mul)[2][2], int g, int t) {
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            float* p = tile + (16 * mt + g) * kHpStage + 8 * nt + 2 * t;
            *reinterpret_cast<float2*>(p) = make_float2(c[mt][nt][0] * mul[mt][0], c[mt][nt][1] * mul[mt][0]);
            *reinterpret_cast<float2*>(p + 8 * kHpStage) = make_float2(c[mt][nt][2] * mul[mt][1], c[mt][nt][3] * mul[mt][1]);
        }
}
// staging tile rows [0, L) -> image columns [gcol0, gcol0 + 32), gcol0 a multiple of 32: a lane
// owns one 16-byte unit (8 columns) of a row per pass
__device__ __forceinline__ void hp_write_img(const float* tile, int L, long long row0, int gcol0, const ig::Img& img, int lane) {
    const int r8 = lane >> 2, u = lane & 3;
    const int gg = (gcol0 >> 3) + u;
    const long long cbase = (long long)(gg >> 3) * img.chunk_stride;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int l = r8 + 8 * i;
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
            *reinterpret_cast<uint4*>(img.lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
template <int TERMS>
__global__ void __launch_bounds__(kHpFwdWarps * 32, 2) attn_hp_fwd_kernel(const AttnArgs a, long long n_items) {
    extern __shared__ __align__(16) float smem[];
    uint8_t* sm = reinterpret_cast<uint8_t*>(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long stride = (long long)gridDim.x * kHpFwdWarps;
    for (long long item = (long long)blockIdx.x * kHpFwdWarps + warp; item < n_items; item += stride) {
    const int L = a.L, D = a.D, dk = a.dk;
    const long long seq = item / a.n_heads;
    const int h = (int)(item - seq * a.n_heads);
    uint8_t* Kb = sm + (size_t)warp * 2 * kHpPairB;       // K pair, later the fp32 output staging
    const uint32_t Ks = (uint32_t)__cvta_generic_to_shared(Kb), Vs = Ks + kHpPairB;
    uint8_t* smask = sm + (size_t)kHpFwdWarps * 2 * kHpPairB + warp * kTile * 8;
    const long long row0 = seq * L;
    const int col = h * dk;
    const int g = lane >> 2, t = lane & 3;

    hp_load_pair<TERMS == 3>(Ks, a.qkv_hi, a.qkv_lo, hp_block_off(seq, 1, h, a.n_heads), L, lane);
    hp_load_pair<TERMS == 3>(Vs, a.qkv_hi, a.qkv_lo, hp_block_off(seq, 2, h, a.n_heads), L, lane);
    const long long qblk = hp_block_off(seq, 0, h, a.n_heads);
    // Q is only ever an A operand: its fragments come straight from the global planes
    uint32_t qh[2][2][4], ql[2][2][4];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = 16 * mt + 8 * (i & 1) + g, d = 16 * ks + 8 * (i >> 1) + 2 * t;
                const long long off = qblk + r * 32 + d;
                qh[ks][mt][i] = r < L ? __ldg(reinterpret_cast<const uint32_t*>(a.qkv_hi + off)) : 0u;
                ql[ks][mt][i] = (TERMS == 3 && r < L) ? __ldg(reinterpret_cast<const uint32_t*>(a.qkv_lo + off)) : 0u;
            }
    const int g0 = col >> 3;
    const bool drop = a.drop.enabled();
    if (drop) {
        // only the ng (<= 5) eight-column groups this head touches: L*ng Philox calls over the warp
        const int ng = ((col + dk + 7) >> 3) - g0;
        for (int it = lane; it < L * ng; it += 32) {
            const int l = it / ng, gi = it - l * ng;
            const uint32_t keep = a.drop.keep8(kDropContext, (uint64_t)(row0 + l), (uint32_t)(g0 + gi));
            smask[l * 8 + gi] = (uint8_t)keep;
            if (a.cmask && g0 + gi < a.mask_bytes) a.cmask[(row0 + l) * a.mask_bytes + g0 + gi] = (uint8_t)keep;
        }
    }
    cp_async_wait_all();
    __syncwarp();
    if (item + stride < n_items) {
        const long long nseq = (item + stride) / a.n_heads;
        hp_prefetch_blocks<TERMS == 3>(a, nseq, (int)(item + stride - nseq * a.n_heads), 0, lane);
    }

    float s[2][4][4];
    zero_frag(s);
    hp_mma<TERMS, false>(s, qh, ql, Ks, lane);
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
    uint32_t ph[2][2][4], pl[2][2][4];
    hp_split_acc<TERMS>(ph, pl, s);             // unnormalised P straight from registers
    float o[2][4][4];
    zero_frag(o);
    hp_mma<TERMS, true>(o, ph, pl, Vs, lane);
    __syncwarp();                               // all reads of K are long done
    float* stage = reinterpret_cast<float*>(Kb);
    hp_stage(stage, o, inv, g, t);              // O = P V / rowsum over the dead K pair
    __syncwarp();
    warp_write_slot<false, kHpStage>(stage, L, dk, row0, col, a.ctx, D, a.ctx_img, drop ? smask : nullptr, g0,
                                     a.drop.scale, nullptr, lane);
    if (a.ctx_img.hi != nullptr)
        pad_image(a.ctx_img, row0, L, D, a.ctx_img.chunks * 64, h == a.n_heads - 1, item == n_items - 1, a.M, lane, true);
    __syncwarp();                               // the staging reads are done before the next item's copies land
    }
}

// dO of one (sequence, head): fp32 rows [row0, row0+L) x columns [col, col+dk) of d_ctx land in a
// 32 x kHpStage fp32 tile by 8-byte cp.async (16 lanes walk the even rows, 16 the odd rows)
__device__ __forceinline__ void hp_request_do(uint32_t tile, const float* d_ctx, long long row0, int D, int col, int L,
                                              int dk, int lane) {
    const int pp = lane & 15, par = lane >> 4;
    if (2 * pp < dk) {
        const float* gp = d_ctx + (row0 + par) * D + col + 2 * pp;
        uint32_t dst = tile + (par * kHpStage + 2 * pp) * 4;
        for (int l = par; l < L; l += 2) {
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(gp) : "memory");
            dst += 2 * kHpStage * 4;
            gp += 2 * D;
        }
    }
}
// the tile -> A fragments (raw fp32 pairs; zero outside [L) x [dk): those cells were never written)
__device__ __forceinline__ void hp_read_do(float2 (&v)[2][2][4], const float* tile, int L, int dk, int g, int t) {
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = 16 * mt + 8 * (i & 1) + g, d = 16 * ks + 8 * (i >> 1) + 2 * t;
                v[ks][mt][i] = (r < L && d < dk) ? *reinterpret_cast<const float2*>(tile + r * kHpStage + d) : make_float2(0.f, 0.f);
            }
}

// ------------------------------------------------------------------------------------------------
// backward
//   P = exp(scale*Q K^T - lse) ; dP = dO V^T ; dS = scale * P o (dP - delta) ; delta = rowsum(P o dP)
//   dV = P^T dO ; dK = dS^T Q ; dQ = dS K        -> d_qkv image, head-padded column order
// ------------------------------------------------------------------------------------------------
template <int TERMS>
__global__ void __launch_bounds__(kHpBwdWarps * 32, 2) attn_hp_bwd_kernel(const AttnArgs a, long long n_items) {
    extern __shared__ __align__(16) float smem[];
    uint8_t* sm = reinterpret_cast<uint8_t*>(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long stride = (long long)gridDim.x * kHpBwdWarps;
    // dO of the NEXT item is requested (cp.async into the dO/dS pair, dead after the dK product) while
    // the current item still has its dQ product and two write-outs to do, and is consumed at the top
    // of the next iteration: its DRAM latency never stalls the warp
    {
        const long long item0 = (long long)blockIdx.x * kHpBwdWarps + warp;
        if (item0 < n_items) {
            const long long seq0 = item0 / a.n_heads;
            const uint32_t G0 = (uint32_t)__cvta_generic_to_shared(sm + (size_t)warp * 4 * kHpPairB) + 3 * kHpPairB;
            hp_request_do(G0, a.d_ctx, seq0 * a.L, a.D, (int)(item0 - seq0 * a.n_heads) * a.dk, a.L, a.dk, lane);
        }
    }
    for (long long item = (long long)blockIdx.x * kHpBwdWarps + warp; item < n_items; item += stride) {
    const int L = a.L, D = a.D, dk = a.dk, DP = 32 * a.n_heads;
    const long long seq = item / a.n_heads;
    const int h = (int)(item - seq * a.n_heads);
    uint8_t* Qb = sm + (size_t)warp * 4 * kHpPairB;
    const uint32_t Qs = (uint32_t)__cvta_generic_to_shared(Qb);   // Q  -> dK staging
    const uint32_t Ks = Qs + kHpPairB;                            // K  -> dQ staging
    const uint32_t Vs = Ks + kHpPairB;                            // V  -> P -> dV staging
    const uint32_t Gs = Vs + kHpPairB;                            // dO -> dS
    const long long row0 = seq * L;
    const int col = h * dk, colp = h * 32;
    const int g = lane >> 2, t = lane & 3;
    const bool drop = a.drop.enabled() && a.cmask != nullptr;

    hp_load_pair<TERMS == 3>(Qs, a.qkv_hi, a.qkv_lo, hp_block_off(seq, 0, h, a.n_heads), L, lane);
    hp_load_pair<TERMS == 3>(Ks, a.qkv_hi, a.qkv_lo, hp_block_off(seq, 1, h, a.n_heads), L, lane);
    hp_load_pair<TERMS == 3>(Vs, a.qkv_hi, a.qkv_lo, hp_block_off(seq, 2, h, a.n_heads), L, lane);
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
    // dO (requested during the previous item; d_ctx carries the context-dropout mask already when the
    // tensor-core data-gradient GEMM produced it): fp32 tile -> fragments, split once, kept as the A
    // operand of dP and stored over the tile as a [row][d] pair for dV's B operand
    uint32_t gh[2][2][4], gl[2][2][4];
    {
        float2 v[2][2][4];
        hp_read_do(v, reinterpret_cast<const float*>(Qb + 3 * kHpPairB), L, dk, g, t);
        __syncwarp();                                   // every lane has its fp32 values before the planes overwrite them
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float2 x = v[ks][mt][i];
                    if (drop) {
                        const int r = 16 * mt + 8 * (i & 1) + g, c = col + 16 * ks + 8 * (i >> 1) + 2 * t;
                        if (r < L && c < col + dk) {
                            const uint32_t keep = (uint32_t)__ldg(a.cmask + (row0 + r) * a.mask_bytes + (c >> 3)) >> (c & 7);
                            x.x = (keep & 1u) ? x.x * a.drop.scale : 0.f;
                            x.y = (keep & 2u) ? x.y * a.drop.scale : 0.f;
                        }
                    }
                    split_pair(x.x, x.y, gh[ks][mt][i], gl[ks][mt][i]);
                }
    }
    hp_store_a<TERMS>(Gs, gh, gl, g, t);
    __syncwarp();
    if (item + stride < n_items) {
        const long long nseq = (item + stride) / a.n_heads;
        const int nh = (int)(item + stride - nseq * a.n_heads);
        hp_prefetch_blocks<TERMS == 3>(a, nseq, nh, 0, lane);
        // the next item's keep bits and log-sum-exps are small, latency-exposed loads: pull them into L2
        if (lane < L) {
            if (drop) prefetch_l2(a.cmask + (nseq * L + lane) * a.mask_bytes + ((nh * dk) >> 3));
            prefetch_l2(a.lse + (nseq * L + lane) * a.n_heads + nh);
        }
    }

    float p[2][4][4], ds[2][4][4];
    zero_frag(p);
    zero_frag(ds);
    {
        uint32_t qh[2][2][4], ql[2][2][4];
        hp_load_a<TERMS, false>(qh, ql, Qs, lane);
        hp_mma<TERMS, false>(p, qh, ql, Ks, lane);      // S
    }
    hp_mma<TERMS, false>(ds, gh, gl, Vs, lane);         // dP
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

    uint32_t sh[2][2][4], sl[2][2][4];                  // dS fragments: A operand of dQ, stored for dK
    hp_split_acc<TERMS>(sh, sl, ds);
    __syncwarp();                                       // all reads of V (dP) are done
    {
        uint32_t ph[2][2][4], pl[2][2][4];
        hp_split_acc<TERMS>(ph, pl, p);
        hp_store_a<TERMS>(Vs, ph, pl, g, t);            // P[row][key] over V
    }
    __syncwarp();
    const float one[2][2] = {{1.f, 1.f}, {1.f, 1.f}};
    const ig::Img& im = a.d_qkv_img;
    float acc[2][4][4];
    uint32_t ah[2][2][4], al[2][2][4];
    zero_frag(acc);
    hp_load_a<TERMS, true>(ah, al, Vs, lane);           // P^T
    hp_mma<TERMS, true>(acc, ah, al, Gs, lane);         // dV[key][d] = sum_row P[row][key] dO[row][d]
    __syncwarp();                                       // all reads of P and dO are done
    hp_stage(reinterpret_cast<float*>(Qb + 2 * kHpPairB), acc, one, g, t);   // dV over P
    hp_store_a<TERMS>(Gs, sh, sl, g, t);                // dS[row][key] over dO
    __syncwarp();
    hp_write_img(reinterpret_cast<const float*>(Qb + 2 * kHpPairB), L, row0, 2 * DP + colp, im, lane);
    zero_frag(acc);
    hp_load_a<TERMS, true>(ah, al, Gs, lane);           // dS^T
    hp_mma<TERMS, true>(acc, ah, al, Qs, lane);         // dK[key][d] = sum_row dS[row][key] Q[row][d]
    __syncwarp();                                       // all reads of Q and of dS^T are done
    if (item + stride < n_items) {                      // the dO/dS pair is dead: the next item's dO lands there
        const long long nseq = (item + stride) / a.n_heads;
        hp_request_do(Gs, a.d_ctx, nseq * L, D, (int)(item + stride - nseq * a.n_heads) * dk, L, dk, lane);
    }
    hp_stage(reinterpret_cast<float*>(Qb), acc, one, g, t);                  // dK over Q
    __syncwarp();
    hp_write_img(reinterpret_cast<const float*>(Qb), L, row0, DP + colp, im, lane);
    zero_frag(acc);
    hp_mma<TERMS, true>(acc, sh, sl, Ks, lane);         // dQ[row][d] = sum_key dS[row][key] K[key][d]
    __syncwarp();                                       // all reads of K are done
    hp_stage(reinterpret_cast<float*>(Qb + kHpPairB), acc, one, g, t);       // dQ over K
    __syncwarp();
    hp_write_img(reinterpret_cast<const float*>(Qb + kHpPairB), L, row0, colp, im, lane);
    // rows [M, rows_pad) of the image are the zero tail of the weight-gradient GEMM's k range
    pad_image(im, row0, L, 0, 0, false, item == n_items - 1, a.M, lane);
    __syncwarp();                                       // the staging reads are done before the next item's copies land
    }
}

// ================================================================================================
// Two warps per (sequence, head)  ("hp2")
// ================================================================================================
// Shared memory holds ten items per SM whatever the kernel does (four pairs = 20 KB each), so with one
// warp per item an SM runs ten warps and the kernel is bound by the latency of each warp's serial
// instruction stream.  Here a PAIR of warps owns an item: warp `mh` owns query rows [16mh, 16mh+16)
// for S, dP, P, dS and dQ, and key rows [16mh, 16mh+16) for dV and dK (A = P^T / dS^T over ALL query
// rows, read from the shared P / dS planes), so no product needs a cross-warp reduction.  Twenty
// warps per SM, half the registers per thread, the item's critical path roughly halved.  The two
// warps meet at a 64-thread named barrier wherever one reads what the other wrote.
constexpr int kHp2Pairs = 5;                   // warp pairs (items in flight) per CTA, 2 CTAs / SM

__device__ __forceinline__ void pair_bar(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

template <int TERMS, bool AT>
__device__ __forceinline__ void hp2_load_a(uint32_t (&ah)[2][4], uint32_t (&al)[2][4], uint32_t pair, int mh, int lane) {
    const int r7 = lane & 7, j0 = (lane >> 3) & 1, j1 = lane >> 4;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
        const uint32_t addr = AT ? pair + (16 * ks + r7 + 8 * j1) * kHpRowB + (16 * mh + 8 * j0) * 2
                                 : pair + (16 * mh + r7 + 8 * j0) * kHpRowB + (16 * ks + 8 * j1) * 2;
        if (AT) ldsm_x4_t(ah[ks], addr); else ldsm_x4(ah[ks], addr);
        if (TERMS == 3) {
            if (AT) ldsm_x4_t(al[ks], addr + kHpPlaneB); else ldsm_x4(al[ks], addr + kHpPlaneB);
        }
    }
}
template <int TERMS>
__device__ __forceinline__ void hp2_split_acc(uint32_t (&ah)[2][4], uint32_t (&al)[2][4], const float (&p)[4][4]) {
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
        split_pair(p[2 * ks][0], p[2 * ks][1], ah[ks][0], al[ks][0]);
        split_pair(p[2 * ks][2], p[2 * ks][3], ah[ks][1], al[ks][1]);
        split_pair(p[2 * ks + 1][0], p[2 * ks + 1][1], ah[ks][2], al[ks][2]);
        split_pair(p[2 * ks + 1][2], p[2 * ks + 1][3], ah[ks][3], al[ks][3]);
    }
}
template <int TERMS>
__device__ __forceinline__ void hp2_store_a(uint32_t pair, const uint32_t (&ah)[2][4], const uint32_t (&al)[2][4], int mh,
                                            int g, int t) {
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t addr = pair + (16 * mh + g + 8 * (i & 1)) * kHpRowB + (16 * ks + 8 * (i >> 1) + 2 * t) * 2;
            sts32(addr, ah[ks][i]);
            if (TERMS == 3) sts32(addr + kHpPlaneB, al[ks][i]);
        }
}
// c[nt] += A[16 x 32] * B[32 x 32]: A fragments in registers, B in a pair ([n][k] or, BT, [k][n])
template <int TERMS, bool BT>
__device__ __forceinline__ void hp2_mma(float (&c)[4][4], const uint32_t (&ah)[2][4], const uint32_t (&al)[2][4], uint32_t pair,
                                        int lane) {
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
        const uint32_t addr = BT ? pair + lane * kHpRowB + nt * 16 : pair + (8 * nt + (lane & 7)) * kHpRowB + (lane >> 3) * 16;
        uint32_t bh[4], bl[4];
        if (BT) ldsm_x4_t(bh, addr); else ldsm_x4(bh, addr);
        if (TERMS == 3) {
            if (BT) ldsm_x4_t(bl, addr + kHpPlaneB); else ldsm_x4(bl, addr + kHpPlaneB);
        }
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            if (TERMS == 3) {
                mma_bf16(c[nt], al[ks], bh[2 * ks], bh[2 * ks + 1]);
                mma_bf16(c[nt], ah[ks], bl[2 * ks], bl[2 * ks + 1]);
            }
            mma_bf16(c[nt], ah[ks], bh[2 * ks], bh[2 * ks + 1]);
        }
    }
}
__device__ __forceinline__ void hp2_zero(float (&c)[4][4]) {
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) c[nt][i] = 0.f;
}
__device__ __forceinline__ void hp2_stage(float* tile, const float (&c)[4][4], int mh, int g, int t) {
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
        float* p = tile + (16 * mh + g) * kHpStage + 8 * nt + 2 * t;
        *reinterpret_cast<float2*>(p) = make_float2(c[nt][0], c[nt][1]);
        *reinterpret_cast<float2*>(p + 8 * kHpStage) = make_float2(c[nt][2], c[nt][3]);
    }
}
// rows [16mh, 16mh+16) of the staging tile -> image columns [gcol0, gcol0 + 32)
__device__ __forceinline__ void hp2_write_img(const float* tile, int mh, int L, long long row0, int gcol0, const ig::Img& img,
                                              int lane) {
    const int r8 = lane >> 2, u = lane & 3;
    const int gg = (gcol0 >> 3) + u;
    const long long cbase = (long long)(gg >> 3) * img.chunk_stride;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int l = 16 * mh + r8 + 8 * i;
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
            *reinterpret_cast<uint4*>(img.lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
    }
}

template <int TERMS>
__global__ void __launch_bounds__(kHp2Pairs * 64, 2) attn_hp2_bwd_kernel(const AttnArgs a, long long n_items) {
    extern __shared__ __align__(16) float smem[];
    uint8_t* sm = reinterpret_cast<uint8_t*>(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pr = warp >> 1, mh = warp & 1, bar = 1 + pr;
    const int g = lane >> 2, t = lane & 3;
    const int L = a.L, D = a.D, dk = a.dk, DP = 32 * a.n_heads;
    const long long stride = (long long)gridDim.x * kHp2Pairs;
    uint8_t* Qb = sm + (size_t)pr * 4 * kHpPairB;
    const uint32_t Qs = (uint32_t)__cvta_generic_to_shared(Qb);   // Q  -> dK staging
    const uint32_t Ks = Qs + kHpPairB;                            // K  -> dQ staging
    const uint32_t Vs = Ks + kHpPairB;                            // V  -> P -> dV staging
    const uint32_t Gs = Vs + kHpPairB;                            // dO (fp32 tile, then planes) -> dS
    float* const stageQ = reinterpret_cast<float*>(Qb);
    float* const stageK = reinterpret_cast<float*>(Qb + kHpPairB);
    float* const stageV = reinterpret_cast<float*>(Qb + 2 * kHpPairB);
    const float* const tileG = reinterpret_cast<const float*>(Qb + 3 * kHpPairB);
    const bool drop = a.drop.enabled() && a.cmask != nullptr;
    const ig::Img& im = a.d_qkv_img;
    {
        const long long item0 = (long long)blockIdx.x * kHp2Pairs + pr;
        if (item0 < n_items && mh == 1) {
            const long long seq0 = item0 / a.n_heads;
            hp_request_do(Gs, a.d_ctx, seq0 * L, D, (int)(item0 - seq0 * a.n_heads) * dk, L, dk, lane);
        }
    }
    for (long long item = (long long)blockIdx.x * kHp2Pairs + pr; item < n_items; item += stride) {
        const long long seq = item / a.n_heads;
        const int h = (int)(item - seq * a.n_heads);
        const long long row0 = seq * L;
        const int col = h * dk, colp = h * 32;
        const bool has_next = item + stride < n_items;
        const long long nseq = (item + stride) / a.n_heads;
        const int nh = (int)(item + stride - nseq * a.n_heads);
        if (mh == 0) {
            hp_load_pair<TERMS == 3>(Qs, a.qkv_hi, a.qkv_lo, hp_block_off(seq, 0, h, a.n_heads), L, lane);
            hp_load_pair<TERMS == 3>(Ks, a.qkv_hi, a.qkv_lo, hp_block_off(seq, 1, h, a.n_heads), L, lane);
        } else {
            hp_load_pair<TERMS == 3>(Vs, a.qkv_hi, a.qkv_lo, hp_block_off(seq, 2, h, a.n_heads), L, lane);
        }
        float lse[2];
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
            const int r = 16 * mh + 8 * hf + g;
            lse[hf] = r < L ? a.lse[(row0 + r) * a.n_heads + h] : 0.f;
        }
        cp_async_wait_all();
        pair_bar(bar);                                      // Q K V and the dO tile are visible to both warps
        uint32_t gh[2][4], gl[2][4];
        {
            float2 v[2][4];
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int r = 16 * mh + 8 * (i & 1) + g, d = 16 * ks + 8 * (i >> 1) + 2 * t;
                    v[ks][i] = (r < L && d < dk) ? *reinterpret_cast<const float2*>(tileG + r * kHpStage + d) : make_float2(0.f, 0.f);
                    if (drop && r < L && d < dk) {
                        const int c = col + d;
                        const uint32_t keep = (uint32_t)__ldg(a.cmask + (row0 + r) * a.mask_bytes + (c >> 3)) >> (c & 7);
                        v[ks][i].x = (keep & 1u) ? v[ks][i].x * a.drop.scale : 0.f;
                        v[ks][i].y = (keep & 2u) ? v[ks][i].y * a.drop.scale : 0.f;
                    }
                }
            pair_bar(bar);                                  // both warps hold their fp32 dO before the planes overwrite the tile
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
#pragma unroll
                for (int i = 0; i < 4; ++i) split_pair(v[ks][i].x, v[ks][i].y, gh[ks][i], gl[ks][i]);
        }
        hp2_store_a<TERMS>(Gs, gh, gl, mh, g, t);           // own rows of the dO planes (dV's B operand)
        if (has_next && mh == 0) {
            hp_prefetch_blocks<TERMS == 3>(a, nseq, nh, 0, lane);
            if (lane < L) prefetch_l2(a.lse + (nseq * L + lane) * a.n_heads + nh);
        }
        float p[4][4], ds[4][4];
        hp2_zero(p);
        hp2_zero(ds);
        {
            uint32_t qh[2][4], ql[2][4];
            hp2_load_a<TERMS, false>(qh, ql, Qs, mh, lane);
            hp2_mma<TERMS, false>(p, qh, ql, Ks, lane);     // S  (own query rows x all keys)
        }
        hp2_mma<TERMS, false>(ds, gh, gl, Vs, lane);        // dP
        float delta[2] = {0.f, 0.f};
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int hf = i >> 1;
                const bool ok = (16 * mh + 8 * hf + g < L) && (8 * nt + 2 * t + (i & 1) < L);
                const float pv = ok ? __expf(p[nt][i] * a.scale - lse[hf]) : 0.f;
                p[nt][i] = pv;
                delta[hf] = fmaf(pv, ds[nt][i], delta[hf]);
            }
        delta[0] = quad_sum(delta[0]);
        delta[1] = quad_sum(delta[1]);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i) ds[nt][i] = p[nt][i] * (ds[nt][i] - delta[i >> 1]) * a.scale;
        uint32_t sh[2][4], sl[2][4];                        // dS fragments: A operand of dQ, stored for dK
        hp2_split_acc<TERMS>(sh, sl, ds);
        {
            uint32_t ph[2][4], pl[2][4];
            hp2_split_acc<TERMS>(ph, pl, p);
            pair_bar(bar);                                  // both warps are done reading V (dP); dO planes complete
            hp2_store_a<TERMS>(Vs, ph, pl, mh, g, t);       // own rows of P[row][key] over V
        }
        pair_bar(bar);                                      // P complete
        float acc[4][4];
        uint32_t ah[2][4], al[2][4];
        hp2_zero(acc);
        hp2_load_a<TERMS, true>(ah, al, Vs, mh, lane);      // P^T, own keys x all query rows
        hp2_mma<TERMS, true>(acc, ah, al, Gs, lane);        // dV[own keys][d]
        pair_bar(bar);                                      // both warps are done reading P and dO
        hp2_stage(stageV, acc, mh, g, t);                   // own rows of dV over P
        hp2_store_a<TERMS>(Gs, sh, sl, mh, g, t);           // own rows of dS[row][key] over dO
        pair_bar(bar);                                      // dS complete (and the own staging rows are visible)
        hp2_write_img(stageV, mh, L, row0, 2 * DP + colp, im, lane);
        hp2_zero(acc);
        hp2_load_a<TERMS, true>(ah, al, Gs, mh, lane);      // dS^T, own keys x all query rows
        hp2_mma<TERMS, true>(acc, ah, al, Qs, lane);        // dK[own keys][d]
        pair_bar(bar);                                      // both warps are done reading Q and dS
        if (has_next && mh == 1) hp_request_do(Gs, a.d_ctx, nseq * L, D, nh * dk, L, dk, lane);   // next item's dO
        hp2_stage(stageQ, acc, mh, g, t);                   // own rows of dK over Q
        __syncwarp();
        hp2_write_img(stageQ, mh, L, row0, DP + colp, im, lane);
        hp2_zero(acc);
        hp2_mma<TERMS, true>(acc, sh, sl, Ks, lane);        // dQ[own rows][d]
        pair_bar(bar);                                      // both warps are done reading K
        hp2_stage(stageK, acc, mh, g, t);                   // own rows of dQ over K
        __syncwarp();
        hp2_write_img(stageK, mh, L, row0, colp, im, lane);
        if (mh == 0) pad_image(im, row0, L, 0, 0, false, item == n_items - 1, a.M, lane);
        pair_bar(bar);                                      // all staging reads are done before the next item's copies land
    }
}

}  // namespace nrms
