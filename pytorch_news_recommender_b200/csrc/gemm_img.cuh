// gemm_img.cuh — tcgen05 GEMMs of the NRMS path over "split-bf16 images".
//
// Every dense contraction of the path (Q|K|V and additive-attention projections, their data
// gradients and their weight gradients; reference nrms_v0.py:53-58,108 and autograd) is
//
//     C[M,N] = epilogue( sum_k A(m,k) * B(n,k) )          fp32 in, fp32 out
//
// computed fp32-grade on the 5th-gen tensor cores through a 3-term bf16 split ("bf16x3"):
// x = hi + lo, hi = bf16(x), lo = bf16(x - hi); A*B ~= Ahi*Bhi + Alo*Bhi + Ahi*Blo with fp32
// accumulation in TMEM (dropped term ~2^-16 relative).
//
// Operand format — the split-bf16 IMAGE of a row-major fp32 matrix X[R, C]:
//   two byte planes (hi, lo); plane = C/64 column CHUNKS, each chunk = R_pad rows of 128 bytes
//   (64 bf16), rows grouped in 8-row / 1024-byte swizzle atoms (16-byte unit index XOR row%8):
//       off(r, c) = (c/64)*R_pad*128 + (r/8)*1024 + (r%8)*128 + ((((c%64)/8) ^ (r%8)) * 16) + (c%8)*2
//   This is byte-for-byte the shared-memory layout tcgen05 expects for SWIZZLE_128B, and the
//   SAME bytes serve both operand orientations:
//     * K-major  (k = columns of X): rows [r0,r0+n) of chunk c are n*128 contiguous bytes;
//     * MN-major (k = rows of X)   : rows [k0,k0+64) of chunk c are 8192 contiguous bytes, one
//       64-wide M/N block of the operand.
//   So a weight image W[out,in] feeds the forward (K-major) and the data gradient (MN-major),
//   and an activation image feeds a forward/data-gradient GEMM (K-major) and the weight
//   gradient (MN-major, k = tokens), all with plain cp.async.bulk copies — no conversion, no
//   transposition and no tensor maps in the GEMM kernel.  The kernels that PRODUCE activations
//   (embedding gather, attention, pooling backward) write the images directly.
//
// Kernel structure (persistent, one CTA per SM, warp-specialised):
//   warp 0     one thread streams operand stages with cp.async.bulk into a shared-memory ring,
//   warp 1     one thread issues tcgen05.mma (cta_group::1, kind::f16, M=128) and commits to
//              the stage-empty / accumulator-full mbarriers,
//   warps 2-5  epilogue: tcgen05.ld the accumulator (TMEM lane = output row), apply
//              bias / tanh+dot / accumulate / dropout-mask, store fp32.
// Accumulators are double-buffered in TMEM when 2*N_T <= 512 columns so the epilogue of one
// tile overlaps the main loop of the next.
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"
#include "tcgen05_ptx.cuh"
#include "profiler.cuh"

namespace nrms {
namespace ig {

constexpr int IMG_CH = 64;          // bf16 columns per chunk
constexpr int IMG_ROW_B = 128;      // bytes per image row
constexpr int IMG_BLOCK_B = 8192;   // 64 rows x 128 B: one MN-major 64-wide block per k-chunk
constexpr int A_TILE_B = 16384;     // 128 rows x 128 B (K-major) == 2 blocks (MN-major)

struct Img {
    uint8_t* hi;
    uint8_t* lo;             // nullptr: single-plane image (gemm_mode 2, plain bf16: nobody writes or reads lo)
    long long chunk_stride;  // rows_pad * 128
    int rows_pad;            // multiple of 128
    int chunks;
};
__host__ __device__ inline int img_rows_pad(long long rows) { return (int)(align_up(rows, 128)); }
__host__ __device__ inline int img_chunks(int cols) { return ceil_div(cols, IMG_CH); }
__host__ __device__ inline long long img_plane_bytes(long long rows, int chunks) {
    return (long long)img_rows_pad(rows) * IMG_ROW_B * chunks;
}
// total bytes of an image (hi plane then lo plane), 1024-aligned by construction
__host__ __device__ inline long long img_bytes(long long rows, int chunks) {
    return 2 * img_plane_bytes(rows, chunks);
}
inline Img img_view(void* base, long long rows, int chunks) {
    Img v;
    v.hi = reinterpret_cast<uint8_t*>(base);
    v.lo = v.hi + img_plane_bytes(rows, chunks);
    v.rows_pad = img_rows_pad(rows);
    v.chunk_stride = (long long)v.rows_pad * IMG_ROW_B;
    v.chunks = chunks;
    return v;
}
// byte offset of the 16-byte unit holding columns [8*g, 8*g+8) of row r (g = global 8-group)
__host__ __device__ __forceinline__ long long img_unit_off(long long chunk_stride, long long r, int g) {
    const int chunk = g >> 3, u = g & 7, r7 = (int)(r & 7);
    return (long long)chunk * chunk_stride + (r >> 3) * 1024 + r7 * 128 + ((u ^ r7) << 4);
}
// (x, y) -> packed bf16 pairs: hi = {bf16(x) low half, bf16(y) high half}, lo = the same of the
// residuals x - hi, y - hi (round to nearest even both times): 6 instructions per pair
__device__ __forceinline__ void split2(float x, float y, uint32_t& hi, uint32_t& lo) {
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(y), "f"(x));
    const float hx = __uint_as_float(hi << 16), hy = __uint_as_float(hi & 0xffff0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(y - hy), "f"(x - hx));
}
// split 8 floats and store them as one 16-byte unit into both planes
__device__ __forceinline__ void img_store8(const Img& im, long long r, int g, const float* x) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) split2(x[2 * j], x[2 * j + 1], hi[j], lo[j]);
    const long long off = img_unit_off(im.chunk_stride, r, g);
    *reinterpret_cast<uint4*>(im.hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    if (im.lo) *reinterpret_cast<uint4*>(im.lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}
__device__ __forceinline__ void img_store8_zero(const Img& im, long long r, int g) {
    const long long off = img_unit_off(im.chunk_stride, r, g);
    *reinterpret_cast<uint4*>(im.hi + off) = make_uint4(0u, 0u, 0u, 0u);
    if (im.lo) *reinterpret_cast<uint4*>(im.lo + off) = make_uint4(0u, 0u, 0u, 0u);
}

// ---- head-padded ("HP") row order of the Q|K|V projection ---------------------------------------
// The short-sequence attention kernels (attention_hp.cuh) want every head's d_k columns to start
// on a 64-byte boundary of a bf16 row, so the projection is computed with every head padded to 32
// output columns: padded row/column rp = which*32h + head*32 + d  (which = Q/K/V, d < d_k real,
// d in [d_k,32) zero weight rows and zero bias).  hp_unpad maps rp back to the row of the
// reference's [3D, D] weight block (nrms_v0.py:36-38), or -1 for a padding row.
__host__ __device__ __forceinline__ int hp_unpad(int rp, int D, int dk) {
    const int DP = (D / dk) * 32;
    const int which = rp / DP, rem = rp - which * DP;
    const int head = rem >> 5, d = rem & 31;
    if (which >= 3 || d >= dk) return -1;
    return which * D + head * dk + d;
}

// ---- fp32 row-major matrix -> image (weights; zero padded) -------------------------------------
// hp_dk > 0: image row r holds source row hp_unpad(r, hp_D, hp_dk) (zeros for padding rows)
__global__ void img_pack_kernel(const float* __restrict__ src, int R, int C, int ld, Img im, int hp_D, int hp_dk) {
    const int groups = im.chunks * 8;
    const long long total = (long long)im.rows_pad * groups;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int g = (int)(i % groups);
        const long long r = i / groups;
        const long long sr = hp_dk > 0 ? hp_unpad((int)r, hp_D, hp_dk) : (r < R ? r : -1);
        float x[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = g * 8 + j;
            x[j] = (sr >= 0 && c < C) ? __ldg(src + sr * ld + c) : 0.f;
        }
        img_store8(im, r, g, x);
    }
}
// two weight matrices of an encoder in ONE launch (the packs are launch-latency-sized kernels): the
// second matrix's units follow the first's in the flattened index space
// ones_col1 >= 0: column ones_col1 of the SECOND image is 1.0 in every real row (the weight-gradient GEMM then
// yields the bias gradient in that output column: reduce_wgrad_kernel)
__global__ void img_pack2_kernel(const float* __restrict__ src0, int R0, int C0, int ld0, Img im0, int hp_D, int hp_dk,
                                 const float* __restrict__ src1, int R1, int C1, int ld1, Img im1, int ones_col1) {
    const long long t0 = (long long)im0.rows_pad * im0.chunks * 8, t1 = (long long)im1.rows_pad * im1.chunks * 8;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < t0 + t1;
         i += (long long)gridDim.x * blockDim.x) {
        const bool first = i < t0;
        const Img& im = first ? im0 : im1;
        const long long k = first ? i : i - t0;
        const int groups = im.chunks * 8;
        const int g = (int)(k % groups);
        const long long r = k / groups;
        const long long sr = first ? (hp_dk > 0 ? hp_unpad((int)r, hp_D, hp_dk) : (r < R0 ? r : -1)) : (r < R1 ? r : -1);
        const float* src = first ? src0 : src1;
        const int C = first ? C0 : C1, ld = first ? ld0 : ld1;
        float x[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = g * 8 + j;
            x[j] = (sr >= 0 && c < C) ? __ldg(src + sr * ld + c) : ((!first && sr >= 0 && c == ones_col1) ? 1.f : 0.f);
        }
        img_store8(im, r, g, x);
    }
}
inline cudaError_t img_pack2(const float* src0, int R0, int C0, int ld0, const Img& im0, int hp_D, int hp_dk,
                             const float* src1, int R1, int C1, int ld1, const Img& im1, cudaStream_t s,
                             int ones_col1 = -1) {
    const long long total = (long long)im0.rows_pad * im0.chunks * 8 + (long long)im1.rows_pad * im1.chunks * 8;
    NRMS_LAUNCH("img_pack", s, img_pack2_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, s>>>(
        src0, R0, C0, ld0, im0, hp_D, hp_dk, src1, R1, C1, ld1, im1, ones_col1));
    return cudaGetLastError();
}
inline cudaError_t img_pack(const float* src, int R, int C, int ld, const Img& im, cudaStream_t s, int hp_D = 0,
                            int hp_dk = 0) {
    const long long total = (long long)im.rows_pad * im.chunks * 8;
    NRMS_LAUNCH("img_pack", s,
                img_pack_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, s>>>(src, R, C, ld, im, hp_D, hp_dk));
    return cudaGetLastError();
}

// ---- the GEMM ----------------------------------------------------------------------------------
enum Epi { EPI_BIAS = 0, EPI_TANH_DOT = 1, EPI_ACCUM = 2, EPI_MASK = 3, EPI_PARTIAL = 4, EPI_POOLADD = 5,
           EPI_BIAS_SPLIT = 6,     // bias (head-padded order) + split-bf16 row-major planes out
           EPI_BIAS_SPLIT_HI = 7 };// the same with the hi plane only (plain-bf16 products: nobody reads lo)

struct IgArgs {
    Img A, B;
    float* C;
    int ldc;
    long long c_split_stride;   // EPI_PARTIAL: partial of split s lives at C + s*c_split_stride
    const float* bias;          // EPI_BIAS / EPI_TANH_DOT: [N]
    const float* qv;            // EPI_TANH_DOT: [N]
    float* dot_out;             // EPI_TANH_DOT: [M]  sum_n tanh(.)*qv[n]
    const uint32_t* mask_bits;  // EPI_MASK / EPI_POOLADD: [M, mask_words] keep bits (bit n%32 of word n/32), or NULL
    int mask_words;
    float mask_scale;           // 1/(1-p)
    const float* row_w;         // EPI_POOLADD: [M] pooling weight of each token row
    const float* seq_vec;       // EPI_POOLADD: [M / seq_len, N] upstream gradient of each sequence
    int seq_len;                // EPI_POOLADD: C[m,n] = acc + row_w[m] * seq_vec[m / seq_len, n]
    uint16_t* Chi;              // EPI_BIAS_SPLIT: head-blocked bf16 planes (hi = bf16(x), lo = bf16(x - hi)),
                                //   [M / seq_len][3][heads] blocks of hp_rows rows x 32 columns (attention_hp.cuh)
    uint16_t* Clo;
    int hp_D, hp_dk;            // EPI_BIAS_SPLIT: bias index of output column n = hp_unpad(n, hp_D, hp_dk)
    int hp_rows;                // EPI_BIAS_SPLIT: rows per head block (32, or 64 for sequences of 33..64 tokens)
    int M, N;                   // valid output rows / columns
    int m_tiles, n_tiles;       // work grid (tiles of 128 rows x N_T columns)
    int k_chunks;               // 64-deep k chunks in total
    int k_steps;                // UMMA K=16 steps in total (<= 4*k_chunks)
    int splits;                 // k splits (weight gradient); 1 otherwise
    int terms;                  // 3 = bf16x3 (fp32-grade), 1 = plain bf16 (hi planes only)
};

// loader warp + MMA warp + EIGHT epilogue warps: two per TMEM lane quadrant, each pair splitting the
// tile's 32-column blocks between them (one epilogue warp per quadrant could not keep up with the main
// loop: the split-plane and tanh epilogues are ALU-bound on a single warp per scheduler).
//   * N_T <= 256: two accumulators (TMEM columns [0,N_T) and [256,256+N_T)), the epilogue of a tile
//     overlaps the main loop of the next;
//   * N_T  > 256: two accumulators do not fit 512 columns.  They are placed at columns [0,N_T) and
//     [512-N_T,512) instead and OVERLAP in [512-N_T, N_T): the epilogue drains the overlap first and
//     releases the accumulator then, so the next tile's main loop runs underneath the rest of the drain.
__host__ __device__ constexpr int ig_epi_warps(int) { return 8; }
__host__ __device__ constexpr int ig_threads(int n_t) { return 64 + 32 * ig_epi_warps(n_t); }

// pair = true: CTA pairs (cta_group::2) — a CTA stages its own 128 rows of A and HALF of the B tile
__host__ __device__ constexpr int ig_stage_bytes(int n_t, bool pair = false) {
    return 2 * A_TILE_B + 2 * (pair ? n_t / 2 : n_t) * IMG_ROW_B;
}
// tail after the ring: 256 B of mbarriers + TMEM slot; N_T <= 256 only (the wider variants have no shared
// memory left): N_T floats of bias + N_T floats of query vector (EPI_TANH_DOT) + 128 floats of row-dot
// exchange between the two warps of a quadrant + one 32 x 128-byte staging buffer per epilogue warp
// (16-byte units XOR-swizzled by row: conflict-free without padding) for coalesced stores
constexpr int IG_XPOSE_BYTES = 32 * 128;
// (the query-vector and row-dot areas exist only where they fit: EPI_TANH_DOT runs with N_T = 208)
__host__ __device__ constexpr int ig_epi_floats(int n_t) {
    return n_t > 256 ? 0 : (n_t + 3) / 4 * 4 + (n_t <= 224 ? (n_t + 3) / 4 * 4 + 128 : 0);
}
__host__ __device__ constexpr int ig_tail_bytes(int n_t) {
    return n_t <= 256 ? 256 + 4 * ig_epi_floats(n_t) + ig_epi_warps(n_t) * IG_XPOSE_BYTES : 384;
}
__host__ __device__ constexpr int ig_stages(int n_t, bool pair = false) {
    return (227 * 1024 - 1024 - ig_tail_bytes(n_t)) / ig_stage_bytes(n_t, pair) >= 4
               ? 4 : (227 * 1024 - 1024 - ig_tail_bytes(n_t)) / ig_stage_bytes(n_t, pair);
}
__host__ __device__ constexpr int ig_smem_bytes(int n_t, bool pair = false) {
    return ig_stages(n_t, pair) * ig_stage_bytes(n_t, pair) + 1024 /*alignment slack*/ + ig_tail_bytes(n_t);
}
static_assert(ig_smem_bytes(256) <= 227 * 1024 && ig_smem_bytes(320) <= 227 * 1024 && ig_smem_bytes(208) <= 227 * 1024,
              "shared-memory budget of the GEMM variants");

// fp32-grade tanh for the additive-attention epilogue: 1 - 2 / (exp(2x) + 1) on ex2.approx / rcp.approx
// (absolute error < 3e-7 over the whole range, exact limits +-1; libdevice's tanhf costs ~4x the
// instructions, and this epilogue runs on one warp per scheduler)
__device__ __forceinline__ float tanh_fast(float x) {
    const float e = __expf(2.f * x);
    return 1.f - __fdividef(2.f, e + 1.f);
}

__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;   // MN-major: bytes between 64-wide blocks
    d |= (uint64_t)(1024 >> 4) << 32;                   // SBO: 8 rows x 128 B
    d |= (uint64_t)1 << 46;                             // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                             // SWIZZLE_128B
    return d;
}
// kind::f16 instruction descriptor: D=F32, A=B=BF16; major bits: 0 = K-major, 1 = MN-major
__host__ __device__ constexpr uint32_t make_idesc2(int n, bool a_mn, bool b_mn, int m = 128) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
           ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// PAIR = true (K-major operands only): the kernel runs as clusters of two CTAs on the two SMs of a TPC and
// issues tcgen05.mma.cta_group::2 with M = 256: CTA r of a pair owns token tile 2p + r (its own A rows, its
// own accumulator rows in its own TMEM, its own epilogue) and stages only HALF of the weight tile — a third
// less operand traffic into each SM and stages small enough for a three-deep ring.  The 1-D bulk copies
// signal mbarriers of their own CTA only, so the second CTA's MMA warp does not issue MMAs but RELAYS its
// stage-full events to the leader (remote mbarrier arrive); the leader's tcgen05.commit multicasts the
// stage-empty / accumulator-full events to both CTAs, and the second CTA's epilogue warps release the
// accumulator on the leader's barrier.
template <bool A_MN, bool B_MN, int N_T, int EPI, bool PAIR = false>
__global__ void __launch_bounds__(ig_threads(N_T), 1) ig_gemm_kernel(const IgArgs a) {
    constexpr int STAGES = ig_stages(N_T, PAIR);
    constexpr int STAGE_B = ig_stage_bytes(N_T, PAIR);
    constexpr int B_PLANE_B = (PAIR ? N_T / 2 : N_T) * IMG_ROW_B;   // B rows this CTA stages
    constexpr bool DOUBLE_ACC = 2 * N_T <= 512;   // two disjoint accumulators
    // TMEM column of the second accumulator: disjoint at 256, else as far right as it fits (overlapping
    // the first in [ACC2, N_T))
    constexpr int ACC2 = DOUBLE_ACC ? 256 : 512 - N_T;
    constexpr int N1 = N_T > 256 ? 256 : N_T;  // first / second UMMA of a k-step (N <= 256 each)
    constexpr int N2 = N_T - N1;
    constexpr int EW = ig_epi_warps(N_T);
    constexpr int N_BLK = (N_T + 31) / 32;     // 32-column blocks of a tile
    constexpr int OV_BLK = DOUBLE_ACC ? 0 : (N_T - ACC2) / 32;   // blocks inside the overlap of the two accumulators
    static_assert(STAGES >= 2, "need at least a double buffer");
    static_assert(N_T % 16 == 0 && N1 % 16 == 0 && N2 % 16 == 0, "UMMA N granularity for M=128");
    static_assert(!B_MN || N_T % 64 == 0, "MN-major operands come in 64-wide blocks");
    static_assert(N_T <= 512, "TMEM has 512 columns");
    static_assert(!PAIR || (!A_MN && !B_MN && N1 % 32 == 0 && N2 % 32 == 0), "CTA pairs: K-major operands, N halves of 16 rows");
    static_assert(DOUBLE_ACC || (ACC2 % 32 == 0 && N_T % 32 == 0 && OV_BLK % 2 == 0 && (N_BLK - OV_BLK) % 2 == 0),
                  "overlapping accumulators: whole 32-column blocks, split evenly between the two warps of a quadrant");

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_B);
    // bars: [0,S) hi planes full | [S,2S) empty | [2S,2S+2) acc full | [2S+2,2S+4) acc empty | [2S+4,3S+4) lo planes full
    //       | PAIR, leader only: [3S+4,4S+4) peer's hi planes full | [4S+4,5S+4) peer's lo planes full
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5 * STAGES + 4);
    static_assert(8 * (5 * STAGES + 4) + 4 <= 256, "barrier block");
    float* s_epi = reinterpret_cast<float*>(smem + STAGES * STAGE_B + 256);   // bias | query vector | row dots (N_T <= 256 only)
    constexpr bool XPOSE = N_T <= 256;   // stage the tile through shared memory for full-line stores
    constexpr int NPAD = (N_T + 3) / 4 * 4;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t smem_base = tc::smem_u32(smem);
    const uint32_t bar_base = tc::smem_u32(bars);
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    auto accf_bar = [&](int b) { return bar_base + 8u * (2 * STAGES + b); };
    auto acce_bar = [&](int b) { return bar_base + 8u * (2 * STAGES + 2 + b); };
    auto full_lo_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 4 + s); };
    auto peer_hi_bar = [&](int s) { return bar_base + 8u * (3 * STAGES + 4 + s); };
    auto peer_lo_bar = [&](int s) { return bar_base + 8u * (4 * STAGES + 4 + s); };
    const uint32_t rank = PAIR ? tc::cluster_ctarank() : 0u;
    const int worker = PAIR ? (int)tc::cluster_id_x() : (int)blockIdx.x;
    const int n_workers = PAIR ? (int)tc::cluster_nclusters_x() : (int)gridDim.x;
    const int m_tiles_w = PAIR ? ceil_div(a.m_tiles, 2) : a.m_tiles;     // token tiles per worker step

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            tc::mbar_init(full_bar(s), 1);
            tc::mbar_init(full_lo_bar(s), 1);
            tc::mbar_init(empty_bar(s), 1);
            tc::mbar_init(peer_hi_bar(s), 1);
            tc::mbar_init(peer_lo_bar(s), 1);
        }
        for (int b = 0; b < 2; ++b) {
            tc::mbar_init(accf_bar(b), 1);
            tc::mbar_init(acce_bar(b), PAIR ? 2 * EW : EW);     // PAIR: both CTAs' epilogue warps release on the leader
        }
        tc::fence_barrier_init();
    }
    if (warp == 1) {
        if (PAIR) tc::tmem_alloc_pair<512>(tc::smem_u32(tmem_slot)); else tc::tmem_alloc<512>(tc::smem_u32(tmem_slot));
    }
    tc::tc_fence_before();
    __syncthreads();
    if (PAIR) tc::cluster_sync_all();     // the peer's barriers are initialised before anything arrives on them
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int total_work = m_tiles_w * a.n_tiles * a.splits;
    const int cps = ceil_div(a.k_chunks, a.splits);   // chunks per split

    if (warp == 0) {
        // =============================== loader ================================================
        if (lane == 0) {
            uint32_t it = 0;
            for (int w = worker; w < total_work; w += n_workers) {
                const int tile = w % (m_tiles_w * a.n_tiles), split = w / (m_tiles_w * a.n_tiles);
                // K-major-A GEMMs walk the token tiles from the LAST one down: the producer kernel wrote the
                // activation image in increasing row order, so its tail is what is still in the 126 MB L2
                const int mt_w = A_MN ? tile / a.n_tiles : m_tiles_w - 1 - tile / a.n_tiles, n_tile = tile % a.n_tiles;
                // PAIR: CTA r takes token tile 2p + r; an odd tile count leaves the last pair's second CTA
                // without a tile: it re-loads the last one (its epilogue stores nothing: rows >= M)
                const int m_tile = PAIR ? min(2 * mt_w + (int)rank, a.m_tiles - 1) : mt_w;
                const int c0 = split * cps, c1 = min(a.k_chunks, c0 + cps);
                for (int kc = c0; kc < c1; ++kc, ++it) {
                    const int s = it % STAGES;
                    tc::mbar_wait(empty_bar(s), ((it / STAGES) & 1u) ^ 1u);
                    const uint32_t dst = smem_base + s * STAGE_B;
                    // The hi planes of a stage signal their own barrier: the Ahi*Bhi products of a k-chunk
                    // start when HALF of the stage has landed, while its lo planes are still in flight.
                    const bool lo = a.terms == 3;     // plain-bf16 mode streams the hi planes only
                    constexpr uint32_t HALF_B = A_TILE_B + B_PLANE_B;
                    const uint32_t dstb = dst + 2 * A_TILE_B;
#pragma unroll
                    for (int pl = 0; pl < 2; ++pl) {
                        if (pl == 1 && !lo) break;
                        const uint32_t bar = pl == 0 ? full_bar(s) : full_lo_bar(s);
                        const uint8_t* Ap = pl == 0 ? a.A.hi : a.A.lo;
                        const uint8_t* Bp = pl == 0 ? a.B.hi : a.B.lo;
                        tc::mbar_arrive_expect_tx(bar, HALF_B);
                        if (A_MN) {
#pragma unroll
                            for (int b = 0; b < 2; ++b) {
                                const long long off = (long long)(m_tile * 2 + b) * a.A.chunk_stride + (long long)kc * IMG_BLOCK_B;
                                tc::bulk_g2s(dst + pl * A_TILE_B + b * IMG_BLOCK_B, Ap + off, IMG_BLOCK_B, bar);
                            }
                        } else {
                            const long long off = (long long)kc * a.A.chunk_stride + (long long)m_tile * A_TILE_B;
                            tc::bulk_g2s(dst + pl * A_TILE_B, Ap + off, A_TILE_B, bar);
                        }
                        if (B_MN) {
#pragma unroll
                            for (int b = 0; b < N_T / 64; ++b) {
                                const long long off = (long long)(n_tile * (N_T / 64) + b) * a.B.chunk_stride + (long long)kc * IMG_BLOCK_B;
                                tc::bulk_g2s(dstb + pl * B_PLANE_B + b * IMG_BLOCK_B, Bp + off, IMG_BLOCK_B, bar);
                            }
                        } else if (PAIR) {
                            // this CTA's half of the rows of each of the (one or two) UMMAs of a k-step
                            const long long off = (long long)kc * a.B.chunk_stride + (long long)n_tile * N_T * IMG_ROW_B;
                            tc::bulk_g2s(dstb + pl * B_PLANE_B, Bp + off + (long long)rank * (N1 / 2) * IMG_ROW_B,
                                         (N1 / 2) * IMG_ROW_B, bar);
                            if (N2 > 0)
                                tc::bulk_g2s(dstb + pl * B_PLANE_B + (N1 / 2) * IMG_ROW_B,
                                             Bp + off + (long long)(N1 + rank * (N2 / 2)) * IMG_ROW_B, (N2 / 2) * IMG_ROW_B, bar);
                        } else {
                            const long long off = (long long)kc * a.B.chunk_stride + (long long)n_tile * B_PLANE_B;
                            tc::bulk_g2s(dstb + pl * B_PLANE_B, Bp + off, B_PLANE_B, bar);
                        }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // =============================== MMA issuer ============================================
        if (PAIR && rank != 0) {
            // second CTA of a pair: relay this CTA's stage-full events to the leader, which issues the MMAs
            if (lane == 0) {
                uint32_t it = 0;
                for (int w = worker; w < total_work; w += n_workers) {
                    const int split = w / (m_tiles_w * a.n_tiles);
                    const int c0 = split * cps, c1 = min(a.k_chunks, c0 + cps);
                    for (int kc = c0; kc < c1; ++kc, ++it) {
                        const int s = it % STAGES;
                        tc::mbar_wait(full_bar(s), (it / STAGES) & 1u);
                        tc::mbar_arrive_cluster(tc::mapa_cluster(peer_hi_bar(s), 0));
                        if (a.terms == 3) {
                            tc::mbar_wait(full_lo_bar(s), (it / STAGES) & 1u);
                            tc::mbar_arrive_cluster(tc::mapa_cluster(peer_lo_bar(s), 0));
                        }
                    }
                }
            }
        } else {
            // Every lane runs this loop with warp-uniform values; the elected lane's MMA / commit instructions
            // are the only ones enabled (tc::umma_*_if).
            const uint32_t issue = tc::elect_one();
            constexpr uint32_t idesc1 = make_idesc2(N1, A_MN, B_MN, PAIR ? 256 : 128);
            constexpr uint32_t idesc2 = make_idesc2(N2 > 0 ? N2 : 32, A_MN, B_MN, PAIR ? 256 : 128);
            constexpr uint32_t A_LBO = A_MN ? IMG_BLOCK_B : 16, B_LBO = B_MN ? IMG_BLOCK_B : 16;
            // descriptor arithmetic on the low word only: bits [0,14) = address / 16 (shared memory is < 256 KB,
            // so adding a byte offset / 16 never carries out of the field), bits [16,30) = LBO / 16
            constexpr uint32_t A_STEP = (A_MN ? 2048 : 32) >> 4, B_STEP = (B_MN ? 2048 : 32) >> 4;   // per K=16 step
            constexpr uint32_t A_LO_OFF = A_TILE_B >> 4, B_LO_OFF = B_PLANE_B >> 4;                  // hi plane -> lo plane
            // second UMMA of a k-step covers columns [N1, N_T): its B rows/blocks start here
            constexpr uint32_t B2_OFF = (B_MN ? (N1 / 64) * IMG_BLOCK_B : (PAIR ? N1 / 2 : N1) * IMG_ROW_B) >> 4;
            auto desc = [](uint32_t lo) { return ((uint64_t)0x40004040u << 32) | lo; };   // SBO 1024, version 1, SWIZZLE_128B
            auto mma = [&](uint32_t d, uint32_t a_lo32, uint32_t b_lo32, uint32_t idesc, uint32_t acc) {
                if (PAIR) tc::umma_bf16_pair_if(issue, d, desc(a_lo32), desc(b_lo32), idesc, acc);
                else tc::umma_bf16_if(issue, d, desc(a_lo32), desc(b_lo32), idesc, acc);
            };
            auto commit = [&](uint32_t bar) {
                if (PAIR) tc::umma_commit_pair_if(issue, bar); else tc::umma_commit_if(issue, bar);
            };
            uint32_t it = 0, tile_it = 0;
            for (int w = worker; w < total_work; w += n_workers, ++tile_it) {
                const int split = w / (m_tiles_w * a.n_tiles);
                const int c0 = split * cps, c1 = min(a.k_chunks, c0 + cps);
                const int buf = (int)(tile_it & 1u);
                if (DOUBLE_ACC) {
                    // the epilogue of the tile two back has drained this accumulator
                    tc::mbar_wait(acce_bar(buf), ((tile_it >> 1) & 1u) ^ 1u);
                } else {
                    // overlapping accumulators: the PREVIOUS tile's epilogue has drained the overlap columns (it
                    // drains them first); the rest of that tile lies outside this tile's columns, and the tile
                    // two back was fully drained before the previous tile's epilogue began
                    tc::mbar_wait(acce_bar(0), (tile_it & 1u) ^ 1u);
                }
                tc::tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(buf * ACC2);
                for (int kc = c0; kc < c1; ++kc, ++it) {
                    const int s = it % STAGES;
                    tc::mbar_wait(full_bar(s), (it / STAGES) & 1u);
                    if (PAIR) tc::mbar_wait(peer_hi_bar(s), (it / STAGES) & 1u);
                    tc::tc_fence_after();
                    const uint32_t sa = smem_base + s * STAGE_B;
                    const uint32_t a_hi = ((sa >> 4) & 0x3FFFu) | ((A_LBO >> 4) << 16);
                    const uint32_t b_hi = (((sa + 2 * A_TILE_B) >> 4) & 0x3FFFu) | ((B_LBO >> 4) << 16);
                    const int steps = min(4, a.k_steps - 4 * kc);
                    // hi x hi of every k-step first (needs the hi planes only) ...
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (j < steps) {
                            const uint32_t acc = (kc > c0 || j > 0) ? 1u : 0u;
                            mma(d_tmem, a_hi + j * A_STEP, b_hi + j * B_STEP, idesc1, acc);
                            if (N2 > 0) mma(d_tmem + N1, a_hi + j * A_STEP, b_hi + B2_OFF + j * B_STEP, idesc2, acc);
                        }
                    }
                    // ... then the two cross terms, once the lo planes have landed
                    if (a.terms == 3) {
                        tc::mbar_wait(full_lo_bar(s), (it / STAGES) & 1u);
                        if (PAIR) tc::mbar_wait(peer_lo_bar(s), (it / STAGES) & 1u);
                        tc::tc_fence_after();
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            if (j < steps) {
                                const uint32_t ah = a_hi + j * A_STEP, bh = b_hi + j * B_STEP;
                                mma(d_tmem, ah + A_LO_OFF, bh, idesc1, 1u);
                                mma(d_tmem, ah, bh + B_LO_OFF, idesc1, 1u);
                                if (N2 > 0) {
                                    mma(d_tmem + N1, ah + A_LO_OFF, bh + B2_OFF, idesc2, 1u);
                                    mma(d_tmem + N1, ah, bh + B_LO_OFF + B2_OFF, idesc2, 1u);
                                }
                            }
                        }
                    }
                    commit(empty_bar(s));   // frees the stage (in both CTAs of a pair) when these MMAs retire
                }
                commit(accf_bar(DOUBLE_ACC ? buf : 0));      // accumulator complete
            }
        }
        __syncwarp();
    } else {
        // =============================== epilogue (warps 2..9) =================================
        const int q = warp & 3;                       // TMEM lane quadrant this warp may read
        const int half = (warp - 2) >> 2;             // which of the quadrant's two warps
        const int et = (warp - 2) * 32 + lane;        // 0..255 within the epilogue group
        constexpr bool SPLIT = EPI == EPI_BIAS_SPLIT || EPI == EPI_BIAS_SPLIT_HI;
        constexpr bool SPLIT_LO = EPI == EPI_BIAS_SPLIT;
        static_assert(!(EPI == EPI_BIAS || EPI == EPI_TANH_DOT || SPLIT) || N_T <= 256,
                      "epilogue staging holds 256 columns");
        static_assert(EPI != EPI_TANH_DOT || N_T <= 224, "query-vector / row-dot staging exists for N_T <= 224 only");
        float* s_bias = s_epi;                        // [NPAD]
        float* s_qv = s_epi + NPAD;                   // [NPAD] (EPI_TANH_DOT)
        float* s_dot = s_epi + 2 * NPAD;              // [128]  (EPI_TANH_DOT: second warp's partial row dots)
        // this warp's staging buffer and the bias vector as SHARED-space addresses (explicit ld/st.shared)
        const uint32_t s_xb = tc::smem_u32(s_epi + ig_epi_floats(N_T)) + (uint32_t)(warp - 2) * IG_XPOSE_BYTES;
        const uint32_t s_bias_a = tc::smem_u32(s_bias);
        uint32_t tile_it = 0;
        auto release = [&](uint32_t bar) {      // PAIR: the leader's MMA warp waits for both CTAs' epilogues
            if (PAIR && rank != 0) tc::mbar_arrive_cluster(tc::mapa_cluster(bar, 0)); else tc::mbar_arrive(bar);
        };
        for (int w = worker; w < total_work; w += n_workers, ++tile_it) {
            const int tile = w % (m_tiles_w * a.n_tiles), split = w / (m_tiles_w * a.n_tiles);
            const int mt_w = A_MN ? tile / a.n_tiles : m_tiles_w - 1 - tile / a.n_tiles, n_tile = tile % a.n_tiles;
            const int m_tile = PAIR ? 2 * mt_w + (int)rank : mt_w;      // (may be one past the last tile: rows >= M)
            const int n0 = n_tile * N_T;
            const int buf = (int)(tile_it & 1u);
            if (EPI == EPI_BIAS || EPI == EPI_TANH_DOT || SPLIT) {
                asm volatile("bar.sync 1, 256;" ::: "memory");     // previous tile's readers are done
                for (int i = et; i < N_T; i += 256) {
                    const int n = n0 + i;
                    if (SPLIT) {
                        const int ns = n < a.N ? hp_unpad(n, a.hp_D, a.hp_dk) : -1;
                        s_bias[i] = ns >= 0 ? __ldg(a.bias + ns) : 0.f;
                        continue;
                    }
                    s_bias[i] = n < a.N ? __ldg(a.bias + n) : 0.f;
                    if (EPI == EPI_TANH_DOT) s_qv[i] = n < a.N ? __ldg(a.qv + n) : 0.f;
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");
            }
            const int m = m_tile * 128 + q * 32 + lane;
            const bool row_ok = m < a.M;
            float rw = 0.f;
            const float* svec = nullptr;
            if (EPI == EPI_POOLADD && row_ok) {
                rw = __ldg(a.row_w + m);
                svec = a.seq_vec + (long long)(m / a.seq_len) * a.N + n0;
            }
            float* cbase = a.C + (EPI == EPI_PARTIAL ? (long long)split * a.c_split_stride : 0ll);
            float* crow = cbase + (long long)m * a.ldc + n0;
            // EPI_BIAS_SPLIT: element offset of the four block rows this lane stores (row 8i + lane%8 of the
            // warp's 32), without the head part: ((seq * 3 * heads) * rows_per_block + l) * 32; -1 past M
            long long split_row_off[4] = {-1, -1, -1, -1};
            if (SPLIT) {
                const int nh = a.hp_D / a.hp_dk;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int mm = m_tile * 128 + q * 32 + 8 * i + (lane & 7);
                    if (mm < a.M) {
                        const int sq = mm / a.seq_len, l = mm - sq * a.seq_len;
                        split_row_off[i] = ((long long)sq * 3 * nh * a.hp_rows + l) * 32;
                    }
                }
            }
            tc::mbar_wait(accf_bar(DOUBLE_ACC ? buf : 0), (DOUBLE_ACC ? (tile_it >> 1) : tile_it) & 1u);
            tc::tc_fence_after();
            const uint32_t t_row = tmem_base + (uint32_t)(buf * ACC2) + ((uint32_t)(q * 32) << 16);
            float dot = 0.f;
            // Block schedule of this warp.  Disjoint accumulators: its share of the blocks in order.
            // Overlapping accumulators: first its share of the OVERLAP blocks (local blocks [N_BLK-OV_BLK, N_BLK)
            // of an even tile = TMEM columns [ACC2, N_T); local blocks [0, OV_BLK) of an odd tile), then the
            // accumulator is released, then its share of the rest.
            constexpr int H0 = (N_BLK + 1) / 2;                 // blocks of the first warp (disjoint case)
            constexpr int MY_OV = OV_BLK / 2, MY_REST = (N_BLK - OV_BLK) / 2;
            static_assert(DOUBLE_ACC || MY_REST >= 1, "the release point sits before the first non-overlap block");
            const int n_mine = DOUBLE_ACC ? (half == 0 ? H0 : N_BLK - H0) : MY_OV + MY_REST;
            auto block_of = [&](int bi) {
                if (DOUBLE_ACC) return half == 0 ? bi : H0 + bi;
                const int ov0 = buf == 0 ? N_BLK - OV_BLK : 0;      // first overlap block (local numbering)
                const int rest0 = buf == 0 ? 0 : OV_BLK;            // first block outside the overlap
                return bi < MY_OV ? ov0 + half * MY_OV + bi : rest0 + half * MY_REST + (bi - MY_OV);
            };
            // The TMEM read of block bi + 1 is in flight while block bi is processed (tcgen05.ld is asynchronous;
            // the epilogue has one or two warps per scheduler, so an exposed load latency is dead time).
            uint32_t vnext[32];
            if (n_mine > 0) tc::tmem_ld32_issue(t_row + 32u * (uint32_t)block_of(0), vnext);
#pragma unroll 1
            for (int bi = 0; bi < n_mine; ++bi) {
                const int blk = block_of(bi);
                if (!DOUBLE_ACC) {
                    if (bi == MY_OV) {
                        // the overlap columns are in registers / stored: the next tile's main loop may start
                        tc::tc_fence_before();
                        __syncwarp();
                        if (lane == 0) release(acce_bar(0));
                    }
                }
                const int cb = 32 * blk;
                // keep bits of this row's 32 columns [n0 + cb, +32): one word (n0 and cb are multiples of 32),
                // requested before the TMEM read so that its latency hides behind it
                uint32_t mword = 0xffffffffu;
                if ((EPI == EPI_MASK || EPI == EPI_POOLADD) && a.mask_bits && row_ok && ((n0 + cb) >> 5) < a.mask_words)
                    mword = __ldg(a.mask_bits + (long long)m * a.mask_words + ((n0 + cb) >> 5));
                float v[32];
                tc::tmem_ld32_arrive(vnext);
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(vnext[i]);
                if (bi + 1 < n_mine) tc::tmem_ld32_issue(t_row + 32u * (uint32_t)block_of(bi + 1), vnext);
                if (SPLIT) {
                    // TMEM gives a lane one output row; 32 columns = one head of the head-padded order, i.e.
                    // one 64-byte row of a head block per plane.  The rows go through the warp's staging
                    // buffer (row = 128 bytes: 4 hi units, 4 lo units, unit index XOR row%8) so that a store
                    // instruction writes 8 consecutive block rows = 512 contiguous bytes (a direct store would
                    // touch 32 rows x 16 bytes: measured, the epilogue then costs more than the main loop)
                    if (n0 + cb < a.N) {
                        uint32_t hi[16], lo[16];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 b4 = tc::lds128f(s_bias_a + 4u * (uint32_t)(cb + 4 * j));
                            if (SPLIT_LO) {
                                split2(v[4 * j] + b4.x, v[4 * j + 1] + b4.y, hi[2 * j], lo[2 * j]);
                                split2(v[4 * j + 2] + b4.z, v[4 * j + 3] + b4.w, hi[2 * j + 1], lo[2 * j + 1]);
                            } else {
                                asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi[2 * j]) : "f"(v[4 * j + 1] + b4.y), "f"(v[4 * j] + b4.x));
                                asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi[2 * j + 1]) : "f"(v[4 * j + 3] + b4.w), "f"(v[4 * j + 2] + b4.z));
                            }
                        }
                        const int l7 = lane & 7;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            tc::sts128(s_xb + lane * 128 + ((j ^ l7) << 4), hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
                            if (SPLIT_LO)
                                tc::sts128(s_xb + lane * 128 + (((4 + j) ^ l7) << 4), lo[4 * j], lo[4 * j + 1], lo[4 * j + 2],
                                           lo[4 * j + 3]);
                        }
                        __syncwarp();
                        // lane -> (row 8i + lane%8, unit lane/8): the 8 lanes of a shared-memory phase read 8
                        // different rows of one unit = 8 different bank groups
                        const int u = lane >> 3;
                        const long long jpart = (long long)((n0 + cb) >> 5) * a.hp_rows * 32 + u * 8;
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            if (split_row_off[i] >= 0) {
                                const int rr = 8 * i + l7;      // rr % 8 == l7
                                const uint4 h4 = tc::lds128(s_xb + rr * 128 + ((u ^ l7) << 4));
                                *reinterpret_cast<uint4*>(a.Chi + split_row_off[i] + jpart) = h4;
                                if (SPLIT_LO) {
                                    const uint4 l4 = tc::lds128(s_xb + rr * 128 + (((4 + u) ^ l7) << 4));
                                    if (a.Clo) *reinterpret_cast<uint4*>(a.Clo + split_row_off[i] + jpart) = l4;
                                }
                            }
                        }
                        __syncwarp();
                    }
                    continue;
                }
                float4 old[8];
                if (EPI == EPI_ACCUM || EPI == EPI_POOLADD) {
                    // batched, independent loads (issued together, consumed below)
#pragma unroll
                    for (int g = 0; g < 8; ++g) {
                        const int n = cb + 4 * g;
                        const bool ok = row_ok && n < N_T && n0 + n < a.N;
                        if (EPI == EPI_ACCUM)
                            old[g] = ok ? *reinterpret_cast<const float4*>(crow + n) : make_float4(0.f, 0.f, 0.f, 0.f);
                        else
                            old[g] = ok ? __ldg(reinterpret_cast<const float4*>(svec + n)) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    const int n = cb + 4 * g;
                    if (n >= N_T) continue;
                    float4 o = make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
                    if (EPI == EPI_BIAS || EPI == EPI_TANH_DOT) {
                        const float4 b4 = tc::lds128f(s_bias_a + 4u * (uint32_t)n);
                        o.x += b4.x; o.y += b4.y; o.z += b4.z; o.w += b4.w;
                    }
                    if (EPI == EPI_TANH_DOT) {
                        const float4 q4 = tc::lds128f(s_bias_a + 4u * (uint32_t)(NPAD + n));   // s_qv follows s_bias
                        o.x = tanh_fast(o.x); o.y = tanh_fast(o.y); o.z = tanh_fast(o.z); o.w = tanh_fast(o.w);
                        dot = fmaf(o.x, q4.x, dot); dot = fmaf(o.y, q4.y, dot);
                        dot = fmaf(o.z, q4.z, dot); dot = fmaf(o.w, q4.w, dot);
                    }
                    if (EPI == EPI_ACCUM) {
                        o.x += old[g].x; o.y += old[g].y; o.z += old[g].z; o.w += old[g].w;
                    }
                    if (EPI == EPI_POOLADD) {
                        o.x = fmaf(rw, old[g].x, o.x); o.y = fmaf(rw, old[g].y, o.y);
                        o.z = fmaf(rw, old[g].z, o.z); o.w = fmaf(rw, old[g].w, o.w);
                    }
                    if (EPI == EPI_MASK || EPI == EPI_POOLADD) {   // POOLADD: the context-dropout mask of the consumer
                        const uint32_t bits = mword >> (4 * g);
                        o.x = (bits & 1u) ? o.x * a.mask_scale : 0.f;
                        o.y = (bits & 2u) ? o.y * a.mask_scale : 0.f;
                        o.z = (bits & 4u) ? o.z * a.mask_scale : 0.f;
                        o.w = (bits & 8u) ? o.w * a.mask_scale : 0.f;
                    }
                    if (XPOSE) {
                        tc::sts128f(s_xb + lane * 128 + ((g ^ (lane & 7)) << 4), o);
                    } else {
                        if (row_ok && n0 + n < a.N) *reinterpret_cast<float4*>(crow + n) = o;
                    }
                }
                if (XPOSE) {
                    // TMEM gives a lane one ROW; a direct store would touch 32 lines per instruction.
                    // Through the warp's swizzled 32 x 128-byte buffer a store instruction writes 4 rows x 128 B.
                    __syncwarp();
                    const int cc = lane & 7;
                    const int n = cb + 4 * cc;
                    if (n < N_T && n0 + n < a.N) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int rr = 4 * i + (lane >> 3);
                            const int mm = m_tile * 128 + q * 32 + rr;
                            if (mm < a.M) {
                                const float4 v4 = tc::lds128f(s_xb + rr * 128 + ((cc ^ (rr & 7)) << 4));
                                *reinterpret_cast<float4*>(cbase + (long long)mm * a.ldc + n0 + n) = v4;
                            }
                        }
                    }
                    __syncwarp();
                }
            }
            if (EPI == EPI_TANH_DOT) {
                // a row's dot with the query vector = the partials of the quadrant's two warps, added in a fixed order
                if (half == 1) s_dot[q * 32 + lane] = dot;
                asm volatile("bar.sync %0, 64;" ::"r"(2 + q) : "memory");
                if (half == 0 && row_ok) a.dot_out[m] = dot + s_dot[q * 32 + lane];
                asm volatile("bar.sync %0, 64;" ::"r"(2 + q) : "memory");   // s_dot may be rewritten by the next tile
            }
            if (DOUBLE_ACC) {
                // release the accumulator
                tc::tc_fence_before();
                __syncwarp();
                if (lane == 0) release(acce_bar(buf));
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (PAIR) tc::cluster_sync_all();     // no CTA exits (or frees TMEM) while its peer may still signal it
    if (warp == 1) {
        if (PAIR) tc::tmem_dealloc_pair<512>(tmem_base); else tc::tmem_dealloc<512>(tmem_base);
    }
}

template <bool A_MN, bool B_MN, int N_T, int EPI, bool PAIR = false>
inline cudaError_t ig_launch(const IgArgs& a, cudaStream_t s, const char* name) {
    constexpr int smem = ig_smem_bytes(N_T, PAIR);
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(ig_gemm_kernel<A_MN, B_MN, N_T, EPI, PAIR>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    if (PAIR) {
        // clusters of two CTAs (the two SMs of a TPC); one pair per two token tiles
        const int pairs = ceil_div(a.m_tiles, 2) * a.n_tiles * a.splits;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * (pairs < kNumSMs / 2 ? pairs : kNumSMs / 2));
        cfg.blockDim = dim3(ig_threads(N_T));
        cfg.dynamicSmemBytes = smem;
        cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        cudaError_t e = cudaSuccess;
        NRMS_LAUNCH(name, s, (e = cudaLaunchKernelEx(&cfg, ig_gemm_kernel<A_MN, B_MN, N_T, EPI, PAIR>, a)));
        return e != cudaSuccess ? e : cudaGetLastError();
    }
    const int total = a.m_tiles * a.n_tiles * a.splits;
    const int grid = total < kNumSMs ? total : kNumSMs;
    NRMS_LAUNCH(name, s, (ig_gemm_kernel<A_MN, B_MN, N_T, EPI, PAIR><<<grid, ig_threads(N_T), smem, s>>>(a)));
    return cudaGetLastError();
}

}  // namespace ig
}  // namespace nrms
