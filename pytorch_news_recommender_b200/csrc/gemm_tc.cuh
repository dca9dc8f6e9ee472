// gemm_tc.cuh — tcgen05 (5th-gen tensor core) GEMM for the dense projections of the path:
//
//     C[M,N] (+)= epilogue( A[M,K] * B[N,K]^T + bias[N] )        fp32 in, fp32 out
//
// Precision: fp32-grade through a 3-term bf16 split ("bf16x3"): x = hi + lo with
// hi = bf16(x), lo = bf16(x - hi);  A*B ~= Ahi*Bhi + Alo*Bhi + Ahi*Blo, accumulated in fp32
// in TMEM (dropped term Alo*Blo ~ 2^-16 relative).  mode x1 issues only Ahi*Bhi.
//
// Structure (one 128 x N_TILE output tile per CTA, K streamed in 64-element chunks through a
// STAGES-deep shared-memory ring):
//   warps 0-3  A producers: load fp32 rows (optionally GATHERED through an int64 row-id list,
//              optionally with the Philox dropout mask), split to bf16 hi/lo, store into the
//              canonical K-major SWIZZLE_128B UMMA layout; afterwards the same warps run the
//              epilogue (tcgen05.ld -> bias/tanh/dropout/accumulate -> global).
//   warp 4     one elected thread issues tcgen05.mma (cta_group::1, kind::f16, M=128,
//              N=N_TILE, K=16) — 4 k-steps x 3 split terms per chunk — and commits to the
//              stage's "empty" mbarrier / the accumulator-ready mbarrier.
//   warp 5     one thread streams the pre-packed B (weight) images with cp.async.bulk
//              (already bf16 hi/lo in the exact swizzled smem image, see pack kernel).
// The accumulator lives in TMEM (N_TILE fp32 columns x 128 lanes).
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"
#include "gemm_simt.cuh"  // GemmArgs
#include "profiler.cuh"

namespace nrms {
namespace tc {

constexpr int BM = 128;          // UMMA M
constexpr int BK = 64;           // bf16 elements per 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int A_BYTES = BM * BK * 2;  // 16 KB per hi or lo tile
constexpr int NUM_PRODUCER_THREADS = 128;
constexpr int THREADS = 192;

__host__ __device__ constexpr int b_bytes(int n_tile) { return n_tile * BK * 2; }
__host__ __device__ constexpr int stage_bytes(int n_tile) { return 2 * A_BYTES + 2 * b_bytes(n_tile); }
__host__ __device__ constexpr int num_stages(int n_tile) {
    return (220 * 1024 - 2048) / stage_bytes(n_tile) >= 4 ? 4 : (220 * 1024 - 2048) / stage_bytes(n_tile);
}
__host__ __device__ constexpr int tmem_cols(int n_tile) {
    return n_tile <= 32 ? 32 : n_tile <= 64 ? 64 : n_tile <= 128 ? 128 : n_tile <= 256 ? 256 : 512;
}

// ---- PTX wrappers ----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
        "l"(src), "r"(bytes), "r"(bar)
        : "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem),
                 "n"(COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS)
                 : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                          uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
          "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),
          "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),
          "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
// start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) | layout=2 (SW128) [61,64)
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;             // LBO (ignored for swizzled K-major), 16 B
    d |= (uint64_t)(1024 >> 4) << 32;   // SBO: 8 rows x 128 B
    d |= (uint64_t)1 << 46;             // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;             // SWIZZLE_128B
    return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D=F32, A=B=BF16, both K-major
__host__ __device__ constexpr uint32_t make_idesc(int n_tile) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n_tile >> 3) << 17) |
           ((uint32_t)(BM >> 4) << 24);
}

// byte offset of element (row r, k element e in [0,64)) inside a K-major SW128 tile
__host__ __device__ __forceinline__ uint32_t sw128_offset(int r, int e) {
    const int chunk = e >> 3;
    return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((chunk ^ (r & 7)) << 4) + (e & 7) * 2);
}

__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
    hi = __float2bfloat16_rn(x);
    lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

// ---- B operand pre-pack ------------------------------------------------------------------------
// Packed image: for n-tile nt, k-chunk kc: [hi tile n_tile x 128 B | lo tile n_tile x 128 B] in
// the swizzled smem layout, contiguous, so a stage is filled by ONE cp.async.bulk.
// Source element B(n,k) = src[n*ld_n + k*ld_k] (lets the same kernel pack W and W^T).
struct PackArgs {
    const float* src;
    uint8_t* dst;
    int N, K, ld_n, ld_k, n_tile, n_tiles, k_chunks;
};
__global__ void pack_b_kernel(const PackArgs p) {
    const long long total = (long long)p.n_tiles * p.k_chunks * p.n_tile * BK;
    const int blk = b_bytes(p.n_tile);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int e = (int)(i % BK);
        long long t = i / BK;
        const int r = (int)(t % p.n_tile);
        t /= p.n_tile;
        const int kc = (int)(t % p.k_chunks);
        const int nt = (int)(t / p.k_chunks);
        const int n = nt * p.n_tile + r, k = kc * BK + e;
        const float x = (n < p.N && k < p.K) ? p.src[(long long)n * p.ld_n + (long long)k * p.ld_k] : 0.f;
        __nv_bfloat16 hi, lo;
        split_bf16(x, hi, lo);
        uint8_t* base = p.dst + ((long long)nt * p.k_chunks + kc) * 2 * blk;
        const uint32_t off = sw128_offset(r, e);
        *reinterpret_cast<__nv_bfloat16*>(base + off) = hi;
        *reinterpret_cast<__nv_bfloat16*>(base + blk + off) = lo;
    }
}
inline int64_t packed_b_bytes(int N, int K, int n_tile) {
    return (int64_t)ceil_div(N, n_tile) * ceil_div(K, BK) * 2 * b_bytes(n_tile);
}

// ---- main kernel -------------------------------------------------------------------------------
struct TcArgs {
    GemmArgs g;            // A, C, bias, a_rows, M, N, K, lda, ldc, accumulate, epilogue, dropout
    const uint8_t* b_packed;
    int k_chunks;
    int terms;             // 3 = bf16x3 (fp32-grade), 1 = plain bf16
};

template <int N_TILE>
__global__ void __launch_bounds__(THREADS, 1) gemm_tc_kernel(const TcArgs a) {
    constexpr int STAGES = num_stages(N_TILE);
    constexpr int B_BYTES = b_bytes(N_TILE);
    constexpr int STAGE_BYTES = stage_bytes(N_TILE);
    constexpr int TMEM_COLS = tmem_cols(N_TILE);
    static_assert(N_TILE % 16 == 0 && N_TILE >= 16 && N_TILE <= 256, "UMMA N for M=128");
    static_assert(STAGES >= 2, "need a double buffer");

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    // bars[0..S) full, bars[S..2S) empty, bars[2S] accumulator ready; then the TMEM base slot
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);

    const GemmArgs& g = a.g;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * BM;
    const int nt = blockIdx.y;
    const int n0 = nt * N_TILE;
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t bar_base = smem_u32(bars);
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    const uint32_t acc_bar = bar_base + 8u * (2 * STAGES);

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), NUM_PRODUCER_THREADS + 1);
            mbar_init(empty_bar(s), 1);
        }
        mbar_init(acc_bar, 1);
        fence_barrier_init();
    }
    if (warp == 4) tmem_alloc<TMEM_COLS>(smem_u32(tmem_slot));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 4) {
        // =============================== A producers ===========================================
        // thread t handles 16-byte chunk c = t & 7 (8 k elements) of rows (t >> 3) + 16*i
        const int c = threadIdx.x & 7;
        const int rbase = threadIdx.x >> 3;
        const float* rowp[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int m = m0 + rbase + 16 * i;
            if (m < g.M) {
                const long long row = g.a_rows ? (long long)g.a_rows[m] : (long long)m;
                rowp[i] = g.A + row * g.lda;
            } else {
                rowp[i] = nullptr;
            }
        }
        for (int kc = 0; kc < a.k_chunks; ++kc) {
            const int s = kc % STAGES;
            const uint32_t round = (uint32_t)(kc / STAGES);
            mbar_wait(empty_bar(s), (round & 1u) ^ 1u);
            uint8_t* a_hi = smem + s * STAGE_BYTES;
            uint8_t* a_lo = a_hi + A_BYTES;
            const int k = kc * BK + c * 8;
            float4 v0[8], v1[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                v0[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                v1[i] = v0[i];
                if (rowp[i]) {
                    if (k < g.K) v0[i] = __ldg(reinterpret_cast<const float4*>(rowp[i] + k));
                    if (k + 4 < g.K) v1[i] = __ldg(reinterpret_cast<const float4*>(rowp[i] + k + 4));
                }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = rbase + 16 * i;
                if (g.drop_on == 1 && rowp[i]) {
                    const uint64_t e = (uint64_t)(m0 + r) * (uint64_t)g.K + (uint64_t)k;
                    if (k < g.K) {
                        const float4 mk = g.drop.mult4(g.drop_sid, e);
                        v0[i].x *= mk.x; v0[i].y *= mk.y; v0[i].z *= mk.z; v0[i].w *= mk.w;
                    }
                    if (k + 4 < g.K) {
                        const float4 mk = g.drop.mult4(g.drop_sid, e + 4);
                        v1[i].x *= mk.x; v1[i].y *= mk.y; v1[i].z *= mk.z; v1[i].w *= mk.w;
                    }
                }
                const float x[8] = {v0[i].x, v0[i].y, v0[i].z, v0[i].w, v1[i].x, v1[i].y, v1[i].z, v1[i].w};
                uint32_t hi[4], lo[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    __nv_bfloat16 h0, l0, h1, l1;
                    split_bf16(x[2 * j], h0, l0);
                    split_bf16(x[2 * j + 1], h1, l1);
                    hi[j] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
                    lo[j] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
                }
                const uint32_t off = (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4));
                *reinterpret_cast<uint4*>(a_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<uint4*>(a_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            }
            fence_proxy_async();          // generic-proxy smem writes -> visible to the tensor core
            mbar_arrive(full_bar(s));
        }
        // =============================== epilogue ==============================================
        mbar_wait(acc_bar, 0);
        tc_fence_after();
        const int m = m0 + warp * 32 + lane;
        const uint32_t t_row = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
        for (int cb = 0; cb < N_TILE; cb += 32) {
            float v[32];
            tmem_ld32(t_row + (uint32_t)cb, v);
            if (m < g.M) {
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int n = n0 + cb + 4 * q;
                    // N_TILE need not be a multiple of the 32-column TMEM load: columns past the
                    // tile belong to the neighbouring CTA
                    if (cb + 4 * q < N_TILE && n < g.N) {
                        float4 o = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                        if (g.bias) {
                            const float4 bb = __ldg(reinterpret_cast<const float4*>(g.bias + n));
                            o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
                        }
                        if (g.epilogue == 1) {
                            o.x = tanhf(o.x); o.y = tanhf(o.y); o.z = tanhf(o.z); o.w = tanhf(o.w);
                        }
                        if (g.drop_on == 3) {
                            const float4 mk = g.drop.mult4(g.drop_sid, (uint64_t)m * (uint64_t)g.N + (uint64_t)n);
                            o.x *= mk.x; o.y *= mk.y; o.z *= mk.z; o.w *= mk.w;
                        }
                        float4* dst = reinterpret_cast<float4*>(g.C + (long long)m * g.ldc + n);
                        if (g.accumulate) {
                            const float4 old = *dst;
                            o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
                        }
                        *dst = o;
                    }
                }
            }
        }
    } else if (warp == 4) {
        // =============================== MMA issuer ============================================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(N_TILE);
            for (int kc = 0; kc < a.k_chunks; ++kc) {
                const int s = kc % STAGES;
                const uint32_t round = (uint32_t)(kc / STAGES);
                mbar_wait(full_bar(s), round & 1u);
                tc_fence_after();
                const uint32_t sa = smem_base + s * STAGE_BYTES;
                const uint64_t a_hi = make_kmajor_sw128_desc(sa);
                const uint64_t a_lo = make_kmajor_sw128_desc(sa + A_BYTES);
                const uint64_t b_hi = make_kmajor_sw128_desc(sa + 2 * A_BYTES);
                const uint64_t b_lo = make_kmajor_sw128_desc(sa + 2 * A_BYTES + B_BYTES);
#pragma unroll
                for (int j = 0; j < BK / UMMA_K; ++j) {
                    const uint64_t adv = (uint64_t)((j * UMMA_K * 2) >> 4);   // 32 B per k-step
                    umma_bf16(tmem_base, a_hi + adv, b_hi + adv, idesc, (kc | j) ? 1u : 0u);
                    if (a.terms == 3) {
                        umma_bf16(tmem_base, a_lo + adv, b_hi + adv, idesc, 1u);
                        umma_bf16(tmem_base, a_hi + adv, b_lo + adv, idesc, 1u);
                    }
                }
                umma_commit(empty_bar(s));                 // frees the stage when the MMAs retire
            }
            umma_commit(acc_bar);                          // accumulator complete
        }
        __syncwarp();
    } else {
        // =============================== B loader ==============================================
        if (lane == 0) {
            const uint8_t* src = a.b_packed + (long long)nt * a.k_chunks * 2 * B_BYTES;
            for (int kc = 0; kc < a.k_chunks; ++kc) {
                const int s = kc % STAGES;
                const uint32_t round = (uint32_t)(kc / STAGES);
                mbar_wait(empty_bar(s), (round & 1u) ^ 1u);
                mbar_arrive_expect_tx(full_bar(s), 2 * B_BYTES);
                bulk_g2s(smem_base + s * STAGE_BYTES + 2 * A_BYTES, src + (long long)kc * 2 * B_BYTES,
                         2 * B_BYTES, full_bar(s));
            }
        }
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc<TMEM_COLS>(tmem_base);
}

template <int N_TILE>
inline cudaError_t launch_tile(const TcArgs& a, int n_tiles, cudaStream_t s, const char* name) {
    constexpr int smem = num_stages(N_TILE) * stage_bytes(N_TILE) + 2048;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<N_TILE>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    dim3 grid(ceil_div(a.g.M, BM), n_tiles);
    NRMS_LAUNCH(name, s, gemm_tc_kernel<N_TILE><<<grid, THREADS, smem, s>>>(a));
    return cudaGetLastError();
}

// N tile per output width: 900 -> 4 x 240 (960), 300 -> 2 x 160 (320), 200 -> 1 x 208
inline int pick_n_tile(int N) {
    if (N <= 208) return 208;
    if (N <= 320) return 160;
    return 240;
}

inline cudaError_t pack_b(const float* src, int N, int K, int ld_n, int ld_k, uint8_t* dst,
                          cudaStream_t s) {
    PackArgs p;
    p.src = src; p.dst = dst; p.N = N; p.K = K; p.ld_n = ld_n; p.ld_k = ld_k;
    p.n_tile = pick_n_tile(N);
    p.n_tiles = ceil_div(N, p.n_tile);
    p.k_chunks = ceil_div(K, BK);
    const long long total = (long long)p.n_tiles * p.k_chunks * p.n_tile * BK;
    NRMS_LAUNCH("pack_b", s, pack_b_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, s>>>(p));
    return cudaGetLastError();
}

// C = epi(A * B^T) with B given as a packed image of a [N,K] matrix (see pack_b)
inline cudaError_t launch(const GemmArgs& g, const uint8_t* b_packed, int terms, cudaStream_t s,
                          const char* name) {
    TcArgs a;
    a.g = g;
    a.b_packed = b_packed;
    a.k_chunks = ceil_div(g.K, BK);
    a.terms = terms;
    const int n_tile = pick_n_tile(g.N);
    const int n_tiles = ceil_div(g.N, n_tile);
    switch (n_tile) {
        case 208: return launch_tile<208>(a, n_tiles, s, name);
        case 160: return launch_tile<160>(a, n_tiles, s, name);
        default: return launch_tile<240>(a, n_tiles, s, name);
    }
}

}  // namespace tc
}  // namespace nrms
