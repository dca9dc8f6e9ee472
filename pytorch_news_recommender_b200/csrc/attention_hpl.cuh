// attention_hpl.cuh — head-padded attention for LONG sequences (65..256 tokens: the 200-click
// history of BASELINE cfg5) on ldmatrix / mma.sync, key-tiled.  Same math, inputs and outputs as
// attention_hpn.cuh (reference nrms_v0.py:13-23, 46-76, 171-173); replaces the CUDA-core FMA
// kernels of attention.cuh for these lengths in the tensor-core GEMM modes.
//
// One CTA per (sequence, head) item, persistent over items; warp w owns rows [16w, 16w+16) of the
// sequence (NW = ceil(L / 16) warps, up to 16).  All of K and V (forward) or Q, K, V and dO
// (backward) of the item sit in shared memory as split-bf16 pairs (attention_hp.cuh geometry: rows
// of 80 bytes), zero-filled up to a multiple of 64 rows, so no tile ever needs a bounds check on
// its operand rows.
//
// forward : flash-attention style.  A warp walks the keys in tiles of 64 with an online softmax
//           (running max / running sum, O rescaled when the max moves): S never exists beyond a
//           16 x 64 register tile.
// backward: NO P / dS planes in shared memory (they would need 2 x 256 x 256 x 4 bytes).  Instead
//           the two reductions run as two phases that each recompute S and dP for their own tiling:
//             phase A (warp owns QUERY rows): per 32-key tile  S, dP -> P, dS -> dQ += dS K
//             phase B (warp owns KEY rows)  : per 32-query tile S^T = K Q^T, dP^T = V dO^T
//                                             -> P^T, dS^T -> dV += P^T dO ; dK += dS^T Q
//           7 small products instead of 5, no cross-warp reduction, two barriers per item.
//           delta_i = sum_j P_ij dP_ij = dO_i . O_i is taken from the saved context image (O after
//           dropout is ctx * (1 - p) wherever dO is non-zero), so phase A needs no extra pass.
#pragma once
#include "attention_hpn.cuh"

namespace nrms {

constexpr int kHplMaxWarps = 16;   // 256 rows

// shared-memory rows of an operand: a whole number of key / query tiles (64 forward, 32 backward)
__host__ __device__ inline int hpl_rows_s(int L, bool fwd) { return (int)align_up(L, fwd ? 64 : 32); }
__host__ __device__ inline int hpl_warps(int L) { return ceil_div(L, 16); }
__host__ __device__ inline size_t attn_hpl_fwd_smem_bytes(int L) {
    return (size_t)hpl_rows_s(L, true) * (2 * 2 * kHpRowB + 8);    // K, V pairs + keep bytes
}
__host__ __device__ inline size_t attn_hpl_bwd_smem_bytes(int L) {
    return (size_t)hpl_rows_s(L, false) * (4 * 2 * kHpRowB + 8);   // Q, K, V, dO pairs + lse, delta
}

__device__ __forceinline__ void cta_bar() { __syncthreads(); }

// rows [r0, r0 + 16) of an operand block, for every 16-row group this warp is responsible for
template <bool LO>
__device__ __forceinline__ void hpl_load_operand(uint32_t pair, int plane, const uint16_t* hi, const uint16_t* lo, long long blk,
                                                 int L, int rows_s, int w, int nw, int lane) {
    for (int r0 = 16 * w; r0 < rows_s; r0 += 16 * nw) hpn_load_rows<LO>(pair, plane, hi, lo, blk, r0, L, lane);
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
template <int TERMS>
__global__ void __launch_bounds__(kHplMaxWarps * 32, 1) attn_hpl_fwd_kernel(const AttnArgs a, long long n_items, int rows) {
    extern __shared__ __align__(16) float smem[];
    uint8_t* sm = reinterpret_cast<uint8_t*>(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const int m0 = 16 * warp, g = lane >> 2, t = lane & 3;
    const int L = a.L, D = a.D, dk = a.dk;
    const int rows_s = hpl_rows_s(L, true), PLANE = rows_s * kHpRowB, PAIR = 2 * PLANE;
    const uint32_t Ks = (uint32_t)__cvta_generic_to_shared(sm), Vs = Ks + PAIR;
    uint8_t* smask = sm + 2 * PAIR;
    float* const stage = reinterpret_cast<float*>(sm);
    const bool drop = a.drop.enabled();
    const int n_kt = ceil_div(L, 64);
    for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
        const long long seq = item / a.n_heads;
        const int h = (int)(item - seq * a.n_heads);
        const long long row0 = seq * L;
        const int col = h * dk, g0 = col >> 3;
        hpl_load_operand<TERMS == 3>(Ks, PLANE, a.qkv_hi, a.qkv_lo, hp_block_off(seq, 1, h, a.n_heads, rows), L, rows_s, warp, nw, lane);
        hpl_load_operand<TERMS == 3>(Vs, PLANE, a.qkv_hi, a.qkv_lo, hp_block_off(seq, 2, h, a.n_heads, rows), L, rows_s, warp, nw, lane);
        // Q is only ever an A operand: the own rows' fragments come straight from the global planes
        uint32_t qh[2][4], ql[2][4];
        const long long qblk = hp_block_off(seq, 0, h, a.n_heads, rows);
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
            const int ng = ((col + dk + 7) >> 3) - g0;
            for (int it = lane; it < (r_end - m0) * ng; it += 32) {
                const int l = m0 + it / ng, gi = it % ng;
                const uint32_t keep = a.drop.keep8(kDropContext, (uint64_t)(row0 + l), (uint32_t)(g0 + gi));
                smask[l * 8 + gi] = (uint8_t)keep;
                if (a.cmask && g0 + gi < a.mask_bytes) a.cmask[(row0 + l) * a.mask_bytes + g0 + gi] = (uint8_t)keep;
            }
        }
        cp_async_wait_all();
        cta_bar();                                          // K and V complete
        float mrun[2] = {-INFINITY, -INFINITY}, lrun[2] = {0.f, 0.f};
        float o[4][4];
        hpn_zero(o);
        for (int kt = 0; kt < n_kt; ++kt) {
            float s[8][4];
            hpn_zero(s);
            hpn_mma_nk<TERMS, 8>(s, qh, ql, Ks + kt * 64 * kHpRowB, PLANE, lane);
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                float m = -INFINITY;
#pragma unroll
                for (int nt = 0; nt < 8; ++nt)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        float& x = s[nt][2 * hf + e];
                        x = (64 * kt + 8 * nt + 2 * t + e < L) ? x * a.scale : -INFINITY;   // scores / sqrt(d_k); no key >= L
                        m = fmaxf(m, x);
                    }
                m = fmaxf(quad_max(m), mrun[hf]);           // tile 0 always holds key 0: m is finite from here on
                const float corr = __expf(mrun[hf] - m);    // exp(-inf) = 0 on the first tile
                float sum = 0.f;
#pragma unroll
                for (int nt = 0; nt < 8; ++nt)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        float& x = s[nt][2 * hf + e];
                        x = __expf(x - m);
                        sum += x;
                    }
                lrun[hf] = lrun[hf] * corr + quad_sum(sum);
                mrun[hf] = m;
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    o[nt][2 * hf] *= corr;
                    o[nt][2 * hf + 1] *= corr;
                }
            }
            uint32_t ph[4][4], pl[4][4];
            hpn_split_acc<TERMS, 4>(ph, pl, s);             // unnormalised P of this key tile, straight from registers
            hpn_mma_kn<TERMS, 4>(o, ph, pl, Vs + kt * 64 * kHpRowB, PLANE, lane);
        }
        float inv[2];
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
            inv[hf] = 1.f / lrun[hf];
            const int r = m0 + 8 * hf + g;
            if (t == 0 && r < L) a.lse[(row0 + r) * a.n_heads + h] = mrun[hf] + __logf(lrun[hf]);
        }
        cta_bar();                                          // every warp is done reading K: the tile goes over it
        hpn_stage(stage, o, inv[0], inv[1], m0, g, t);      // O = P V / rowsum, own rows
        __syncwarp();
        warp_write_slot<false, kHpStage>(stage, r_end, dk, row0, col, a.ctx, D, a.ctx_img, drop ? smask : nullptr, g0,
                                         a.drop.scale, nullptr, lane, m0);
        if (a.ctx_img.hi != nullptr && warp == 0)
            pad_image(a.ctx_img, row0, L, D, a.ctx_img.chunks * 64, h == a.n_heads - 1, item == n_items - 1, a.M, lane, true);
        cta_bar();                                          // staging reads are done before the next item's copies land
    }
}

// ------------------------------------------------------------------------------------------------
// backward
//   P = exp(scale*Q K^T - lse) ; dP = dO V^T ; dS = scale * P o (dP - delta) ; delta = rowsum(dO o O)
//   dV = P^T dO ; dK = dS^T Q ; dQ = dS K        -> d_qkv image, head-padded column order
// ------------------------------------------------------------------------------------------------
template <int TERMS>
__global__ void __launch_bounds__(kHplMaxWarps * 32, 1) attn_hpl_bwd_kernel(const AttnArgs a, long long n_items, int rows) {
    extern __shared__ __align__(16) float smem[];
    uint8_t* sm = reinterpret_cast<uint8_t*>(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const int m0 = 16 * warp, g = lane >> 2, t = lane & 3;
    const int L = a.L, D = a.D, dk = a.dk, DP = 32 * a.n_heads;
    const int rows_s = hpl_rows_s(L, false), PLANE = rows_s * kHpRowB, PAIR = 2 * PLANE;
    const uint32_t Qs = (uint32_t)__cvta_generic_to_shared(sm), Ks = Qs + PAIR, Vs = Ks + PAIR, Gs = Vs + PAIR;
    float* const stageQ = reinterpret_cast<float*>(sm);               // dK goes over Q
    float* const stageK = reinterpret_cast<float*>(sm + PAIR);        // dQ goes over K
    float* const stageV = reinterpret_cast<float*>(sm + 2 * PAIR);    // dV goes over V
    const float* const tileG = reinterpret_cast<const float*>(sm + 3 * PAIR);
    float* const s_lse = reinterpret_cast<float*>(sm + 4 * PAIR);
    float* const s_delta = s_lse + rows_s;
    const ig::Img& im = a.d_qkv_img;
    const int n_t = ceil_div(L, 32);                                   // 32-wide key / query tiles that hold real rows
    // d_ctx arrives with the context-dropout mask and 1/(1-p) applied (dgrad GEMM epilogue); O = ctx * (1-p)
    const float o_scale = a.drop.enabled() ? 1.f / a.drop.scale : 1.f;
    for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
        const long long seq = item / a.n_heads;
        const int h = (int)(item - seq * a.n_heads);
        const long long row0 = seq * L;
        const int col = h * dk, colp = h * 32;
        hpl_load_operand<TERMS == 3>(Qs, PLANE, a.qkv_hi, a.qkv_lo, hp_block_off(seq, 0, h, a.n_heads, rows), L, rows_s, warp, nw, lane);
        hpl_load_operand<TERMS == 3>(Ks, PLANE, a.qkv_hi, a.qkv_lo, hp_block_off(seq, 1, h, a.n_heads, rows), L, rows_s, warp, nw, lane);
        hpl_load_operand<TERMS == 3>(Vs, PLANE, a.qkv_hi, a.qkv_lo, hp_block_off(seq, 2, h, a.n_heads, rows), L, rows_s, warp, nw, lane);
        hpn_request_do(Gs, a.d_ctx, row0, D, col, m0, L, dk, lane);     // own rows of dO as an fp32 tile
        for (int r = threadIdx.x; r < rows_s; r += blockDim.x) s_lse[r] = r < L ? a.lse[(row0 + r) * a.n_heads + h] : 0.f;
        cp_async_wait_all();
        cta_bar();                                          // Q K V, the dO tile and lse are visible to every warp
        // own rows of dO: fp32 tile -> registers (+ delta from the saved context image) -> split planes
        uint32_t gh[2][4], gl[2][4];
        {
            float2 v[2][4];
            float dl[2] = {0.f, 0.f};
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int r = m0 + 8 * (i & 1) + g, d = 16 * ks + 8 * (i >> 1) + 2 * t;
                    const bool ok = r < L && d < dk;
                    v[ks][i] = ok ? *reinterpret_cast<const float2*>(tileG + r * kHpStage + d) : make_float2(0.f, 0.f);
                    if (ok) {
                        const int c = col + d;
                        const long long off = ig::img_unit_off(a.ctx_img.chunk_stride, row0 + r, c >> 3) + (c & 7) * 2;
                        const uint32_t ch = __ldg(reinterpret_cast<const uint32_t*>(a.ctx_img.hi + off));
                        const uint32_t cl = (TERMS == 3 && a.ctx_img.lo) ? __ldg(reinterpret_cast<const uint32_t*>(a.ctx_img.lo + off)) : 0u;
                        const float ox = __uint_as_float(ch << 16) + __uint_as_float(cl << 16);
                        const float oy = __uint_as_float(ch & 0xffff0000u) + __uint_as_float(cl & 0xffff0000u);
                        dl[i & 1] = fmaf(v[ks][i].x, ox, fmaf(v[ks][i].y, oy, dl[i & 1]));
                    }
                }
            dl[0] = quad_sum(dl[0]) * o_scale;
            dl[1] = quad_sum(dl[1]) * o_scale;
            if (t == 0) {
                s_delta[m0 + g] = dl[0];
                s_delta[m0 + 8 + g] = dl[1];
            }
            cta_bar();                                      // every warp holds its fp32 dO before the planes overwrite the tile
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
#pragma unroll
                for (int i = 0; i < 4; ++i) split_pair(v[ks][i].x, v[ks][i].y, gh[ks][i], gl[ks][i]);
        }
        hpn_store_a<TERMS, 2>(Gs, kHpRowB, PLANE, gh, gl, m0, g, t);
        // rows no warp owns ([16 nw, rows_s)): zero dO planes and zero delta (they are B-operand rows of phase B)
        for (int i = threadIdx.x; i < (rows_s - 16 * nw) * (kHpRowB / 4); i += blockDim.x) {
            const uint32_t addr = Gs + 16 * nw * kHpRowB + 4 * i;
            sts32(addr, 0u);
            if (TERMS == 3) sts32(addr + PLANE, 0u);
        }
        for (int r = 16 * nw + threadIdx.x; r < rows_s; r += blockDim.x) s_delta[r] = 0.f;
        cta_bar();                                          // dO planes and delta complete
        // ---------------- phase A: own QUERY rows, key tiles of 32 -> dQ
        float accq[4][4];
        hpn_zero(accq);
        {
            uint32_t qh[2][4], ql[2][4];
            hpn_load_a<TERMS, false, 2>(qh, ql, Qs, kHpRowB, PLANE, m0, lane);
            const float lse0 = s_lse[m0 + g], lse1 = s_lse[m0 + 8 + g];
            const float dl0 = s_delta[m0 + g], dl1 = s_delta[m0 + 8 + g];
            const bool row_ok0 = m0 + g < L, row_ok1 = m0 + 8 + g < L;
            for (int kt = 0; kt < n_t; ++kt) {
                float p[4][4], ds[4][4];
                hpn_zero(p);
                hpn_zero(ds);
                hpn_mma_nk<TERMS, 4>(p, qh, ql, Ks + kt * 32 * kHpRowB, PLANE, lane);    // S  (own rows x 32 keys)
                hpn_mma_nk<TERMS, 4>(ds, gh, gl, Vs + kt * 32 * kHpRowB, PLANE, lane);   // dP
#pragma unroll
                for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int hf = i >> 1;
                        const bool ok = (hf ? row_ok1 : row_ok0) && (32 * kt + 8 * nt + 2 * t + (i & 1) < L);
                        const float pv = ok ? __expf(p[nt][i] * a.scale - (hf ? lse1 : lse0)) : 0.f;
                        ds[nt][i] = pv * (ds[nt][i] - (hf ? dl1 : dl0)) * a.scale;
                    }
                uint32_t sh[2][4], sl[2][4];
                hpn_split_acc<TERMS, 2>(sh, sl, ds);
                hpn_mma_kn<TERMS, 2>(accq, sh, sl, Ks + kt * 32 * kHpRowB, PLANE, lane);  // dQ += dS K
            }
        }
        // ---------------- phase B: own KEY rows, query tiles of 32 -> dV, dK
        uint32_t kh_[2][4], kl_[2][4], vh_[2][4], vl_[2][4];
        hpn_load_a<TERMS, false, 2>(kh_, kl_, Ks, kHpRowB, PLANE, m0, lane);
        hpn_load_a<TERMS, false, 2>(vh_, vl_, Vs, kHpRowB, PLANE, m0, lane);
        cta_bar();                                          // every warp is done with K and V as shared operands
        hpn_stage(stageK, accq, 1.f, 1.f, m0, g, t);        // own rows of dQ over K
        __syncwarp();
        hpn_write_img(stageK, m0, L, row0, colp, im, lane);
        float acck[4][4], accv[4][4];
        hpn_zero(acck);
        hpn_zero(accv);
        {
            const bool key_ok0 = m0 + g < L, key_ok1 = m0 + 8 + g < L;
            for (int qt = 0; qt < n_t; ++qt) {
                float st[4][4], dpt[4][4];
                hpn_zero(st);
                hpn_zero(dpt);
                hpn_mma_nk<TERMS, 4>(st, kh_, kl_, Qs + qt * 32 * kHpRowB, PLANE, lane);   // S^T  (own keys x 32 queries)
                hpn_mma_nk<TERMS, 4>(dpt, vh_, vl_, Gs + qt * 32 * kHpRowB, PLANE, lane);  // dP^T
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    const int q0 = 32 * qt + 8 * nt + 2 * t;
                    const float2 ls = *reinterpret_cast<const float2*>(s_lse + q0);
                    const float2 dl = *reinterpret_cast<const float2*>(s_delta + q0);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int e = i & 1;
                        const bool ok = ((i >> 1) ? key_ok1 : key_ok0) && (q0 + e < L);
                        const float pv = ok ? __expf(st[nt][i] * a.scale - (e ? ls.y : ls.x)) : 0.f;
                        st[nt][i] = pv;
                        dpt[nt][i] = pv * (dpt[nt][i] - (e ? dl.y : dl.x)) * a.scale;
                    }
                }
                uint32_t ah[2][4], al[2][4];
                hpn_split_acc<TERMS, 2>(ah, al, st);
                hpn_mma_kn<TERMS, 2>(accv, ah, al, Gs + qt * 32 * kHpRowB, PLANE, lane);   // dV += P^T dO
                hpn_split_acc<TERMS, 2>(ah, al, dpt);
                hpn_mma_kn<TERMS, 2>(acck, ah, al, Qs + qt * 32 * kHpRowB, PLANE, lane);   // dK += dS^T Q
            }
        }
        cta_bar();                                          // every warp is done reading Q and dO
        hpn_stage(stageQ, acck, 1.f, 1.f, m0, g, t);        // own rows of dK over Q
        hpn_stage(stageV, accv, 1.f, 1.f, m0, g, t);        // own rows of dV over V
        __syncwarp();
        hpn_write_img(stageQ, m0, L, row0, DP + colp, im, lane);
        hpn_write_img(stageV, m0, L, row0, 2 * DP + colp, im, lane);
        if (warp == 0) pad_image(im, row0, L, 0, 0, false, item == n_items - 1, a.M, lane);
        cta_bar();                                          // staging reads are done before the next item's copies land
    }
}

}  // namespace nrms
