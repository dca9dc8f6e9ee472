// gemm_simt.cuh — fp32 CUDA-core GEMM used for every dense contraction of the path when
// gemm_mode == 0 (exact fp32; the tcgen05 path in gemm_tc.cuh replaces the hot shapes).
//
//   C[M,N] (+)= epilogue( A(m,k) * B(k,n) + bias[n] )
//
// Operand access is templated on which index is contiguous in memory so the same kernel
// serves the three contraction shapes of an nn.Linear:
//   forward   y = x W^T + b   : A K-contiguous, B K-contiguous (W is [N,K])
//   data grad dx = dy W       : A K-contiguous, B N-contiguous (W is [K,N] for this product)
//   weight grad dW = dy^T x   : A M-contiguous (dy^T), B N-contiguous (x), reduction split over
//                               blockIdx.z into partials.
// The embedding gather and both dropouts happen in the producing kernels (gather.cuh,
// attention.cuh); the data gradient w.r.t. the embedding rows applies the stored keep bits.
// All contiguous extents are feature dims (300/900/200) => multiples of 4 => float4 access.
#pragma once
#include "common.cuh"
#include "profiler.cuh"

namespace nrms {

struct GemmArgs {
    const float* A;
    const float* B;
    float* C;
    const float* bias;      // [N] or nullptr
    int M, N, K;
    int lda, ldb, ldc;
    int k_chunk;               // reduction range per blockIdx.z (K when not split)
    long long c_split_stride;  // elements between per-z outputs
    int accumulate;            // C += result instead of C = result
    int epilogue;              // 0 none, 1 tanh
    const uint8_t* mask;       // optional keep bits of C [M, mask_bytes] (byte n/8, bit n%8)
    int mask_bytes;
    float mask_scale;          // 1/(1-p)
};

constexpr int GBM = 128, GBN = 128, GBK = 16, GTHREADS = 256, GPAD = 4;

template <bool A_KC, bool B_KC>
__global__ void __launch_bounds__(GTHREADS, 2) gemm_simt_kernel(const GemmArgs g) {
    __shared__ __align__(16) float As[GBK][GBM + GPAD];
    __shared__ __align__(16) float Bs[GBK][GBN + GPAD];

    const int t = threadIdx.x;
    const int m0 = blockIdx.x * GBM, n0 = blockIdx.y * GBN;
    const int kbeg = blockIdx.z * g.k_chunk;
    const int kend = min(g.K, kbeg + g.k_chunk);
    float* __restrict__ C = g.C + (long long)blockIdx.z * g.c_split_stride;

    // ---- global -> register staging ---------------------------------------------------
    float4 ra[2], rb[2];
    const float* a_ptr[2];
    const float* b_ptr[2];
    int a_i0[2], a_i1[2], b_i0[2], b_i1[2];  // (tile-local indices) see below
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int idx = t + i * GTHREADS;
        if (A_KC) {
            a_i0[i] = idx >> 2;        // m_l
            a_i1[i] = (idx & 3) << 2;  // k_l (x4)
            const int m = m0 + a_i0[i];
            a_ptr[i] = (m < g.M) ? g.A + (long long)m * g.lda : nullptr;
        } else {
            a_i0[i] = idx >> 5;        // k_l
            a_i1[i] = (idx & 31) << 2;  // m_l (x4)
            a_ptr[i] = (m0 + a_i1[i] < g.M) ? g.A + (m0 + a_i1[i]) : nullptr;
        }
        if (B_KC) {
            b_i0[i] = idx >> 2;        // n_l
            b_i1[i] = (idx & 3) << 2;  // k_l (x4)
            const int n = n0 + b_i0[i];
            b_ptr[i] = (n < g.N) ? g.B + (long long)n * g.ldb : nullptr;
        } else {
            b_i0[i] = idx >> 5;         // k_l
            b_i1[i] = (idx & 31) << 2;  // n_l (x4)
            b_ptr[i] = (n0 + b_i1[i] < g.N) ? g.B + (n0 + b_i1[i]) : nullptr;
        }
    }

    auto load_regs = [&](int k0) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (A_KC) {
                const int k = k0 + a_i1[i];
                if (a_ptr[i] && k < kend) v = __ldg(reinterpret_cast<const float4*>(a_ptr[i] + k));
            } else {
                const int k = k0 + a_i0[i];
                if (a_ptr[i] && k < kend)
                    v = __ldg(reinterpret_cast<const float4*>(a_ptr[i] + (long long)k * g.lda));
            }
            ra[i] = v;
            float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
            if (B_KC) {
                const int k = k0 + b_i1[i];
                if (b_ptr[i] && k < kend) w = __ldg(reinterpret_cast<const float4*>(b_ptr[i] + k));
            } else {
                const int k = k0 + b_i0[i];
                if (b_ptr[i] && k < kend)
                    w = __ldg(reinterpret_cast<const float4*>(b_ptr[i] + (long long)k * g.ldb));
            }
            rb[i] = w;
        }
    };
    auto store_smem = [&]() {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            if (A_KC) {
                As[a_i1[i] + 0][a_i0[i]] = ra[i].x;
                As[a_i1[i] + 1][a_i0[i]] = ra[i].y;
                As[a_i1[i] + 2][a_i0[i]] = ra[i].z;
                As[a_i1[i] + 3][a_i0[i]] = ra[i].w;
            } else {
                *reinterpret_cast<float4*>(&As[a_i0[i]][a_i1[i]]) = ra[i];
            }
            if (B_KC) {
                Bs[b_i1[i] + 0][b_i0[i]] = rb[i].x;
                Bs[b_i1[i] + 1][b_i0[i]] = rb[i].y;
                Bs[b_i1[i] + 2][b_i0[i]] = rb[i].z;
                Bs[b_i1[i] + 3][b_i0[i]] = rb[i].w;
            } else {
                *reinterpret_cast<float4*>(&Bs[b_i0[i]][b_i1[i]]) = rb[i];
            }
        }
    };

    // ---- 8x8 micro-tile per thread, split 4+4 at distance 64 (conflict-free LDS.128) ---
    const int ty = t >> 4, tx = t & 15;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    if (kbeg < kend) {
        load_regs(kbeg);
        store_smem();
    }
    __syncthreads();
    for (int k0 = kbeg; k0 < kend; k0 += GBK) {
        const bool more = (k0 + GBK) < kend;
        if (more) load_regs(k0 + GBK);
#pragma unroll
        for (int kk = 0; kk < GBK; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk][64 + tx * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
        if (more) {
            store_smem();
            __syncthreads();
        }
    }

    // ---- epilogue ---------------------------------------------------------------------
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (m >= g.M) continue;
#pragma unroll
        for (int jh = 0; jh < 2; ++jh) {
            const int n = n0 + (jh == 0 ? tx * 4 : 64 + tx * 4);
            if (n >= g.N) continue;
            float4 v = make_float4(acc[i][jh * 4 + 0], acc[i][jh * 4 + 1], acc[i][jh * 4 + 2],
                                   acc[i][jh * 4 + 3]);
            if (g.bias) {
                const float4 bb = __ldg(reinterpret_cast<const float4*>(g.bias + n));
                v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
            }
            if (g.epilogue == 1) {
                v.x = tanhf(v.x); v.y = tanhf(v.y); v.z = tanhf(v.z); v.w = tanhf(v.w);
            }
            if (g.mask) {
                const uint32_t bits = (uint32_t)g.mask[(long long)m * g.mask_bytes + (n >> 3)] >> (n & 7);
                v.x = (bits & 1u) ? v.x * g.mask_scale : 0.f;
                v.y = (bits & 2u) ? v.y * g.mask_scale : 0.f;
                v.z = (bits & 4u) ? v.z * g.mask_scale : 0.f;
                v.w = (bits & 8u) ? v.w * g.mask_scale : 0.f;
            }
            float4* dst = reinterpret_cast<float4*>(C + (long long)m * g.ldc + n);
            if (g.accumulate) {
                const float4 o = *dst;
                v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
            }
            *dst = v;
        }
    }
}

// Host launcher.  Returns cudaGetLastError().
inline cudaError_t launch_gemm_simt(const GemmArgs& g, bool a_kc, bool b_kc, int splits,
                                    cudaStream_t s, const char* name = "gemm_simt") {
    dim3 grid(ceil_div(g.M, GBM), ceil_div(g.N, GBN), splits);
    if (a_kc && b_kc)
        NRMS_LAUNCH(name, s, gemm_simt_kernel<true, true><<<grid, GTHREADS, 0, s>>>(g));
    else if (a_kc && !b_kc)
        NRMS_LAUNCH(name, s, gemm_simt_kernel<true, false><<<grid, GTHREADS, 0, s>>>(g));
    else if (!a_kc && !b_kc)
        NRMS_LAUNCH(name, s, gemm_simt_kernel<false, false><<<grid, GTHREADS, 0, s>>>(g));
    else
        NRMS_LAUNCH(name, s, gemm_simt_kernel<false, true><<<grid, GTHREADS, 0, s>>>(g));
    return cudaGetLastError();
}

}  // namespace nrms
