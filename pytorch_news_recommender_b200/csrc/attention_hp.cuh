// attention_hp.cuh — self-attention for sequences of up to 32 tokens over the HEAD-PADDED
// split-bf16 Q|K|V planes the projection GEMM writes (gemm_img.cuh: EPI_BIAS_SPLIT, hp_unpad):
// the shared pieces and the one-warp-per-(sequence, head) FORWARD kernel.  The backward and the
// 33..64-token kernels are in attention_hpn.cuh.  Same math and outputs as attention_mma.cuh
// (reference nrms_v0.py:13-23, 46-76, 171-173).
//
// What changes against attention_mma.cuh is how operands reach the tensor pipe:
//   * Q|K|V arrive as two bf16 planes (hi = bf16(x), lo = bf16(x - hi)), every head 32 columns wide,
//     head-blocked (one operand = one contiguous 2 KB block per plane): cp.async.cg of 16 bytes with
//     zero-fill for rows >= L, and NO fp32 -> bf16 split inside the kernel;
//   * every fragment is one ldmatrix.x4 (4 per operand plane per product) instead of 16-32 scalar
//     shared-memory loads plus ~7 ALU instructions per pair for the split; operands needed
//     transposed use ldmatrix.trans on the same row-major planes — no transposed scatter stores;
//   * the d_qkv image keeps the padded column order, so a head's output row is 4 aligned 16-byte
//     units per plane: one lane stores a whole unit (8 rows x 64 B per store instruction).
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

__host__ __device__ inline size_t attn_hp_fwd_smem_bytes() {
    return (size_t)kHpFwdWarps * (2 * kHpPairB + kTile * 8);
}

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
__device__ __forceinline__ void ldsm_x2_t(uint32_t (&r)[2], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// element offset of the 32 x 32 block of (sequence, which = Q/K/V, head) in a plane: the planes are
// HEAD-BLOCKED, [n_seq][3][h] blocks of 32 (64 for sequences of 33..64 tokens) rows x 32 bf16
// (2 KB contiguous per operand), so that a
// warp's operand is one contiguous 2 KB read instead of 30 64-byte pieces at a 1,920-byte stride
__host__ __device__ __forceinline__ long long hp_block_off(long long seq, int which, int head, int n_heads, int rows = 32) {
    return ((seq * 3 + which) * n_heads + head) * (long long)(rows * 32);
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

}  // namespace nrms
