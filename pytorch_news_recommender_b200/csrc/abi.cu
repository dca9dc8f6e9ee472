// abi.cu — the extern "C" surface of libnrms_b200.so (see include/nrms_b200.h) and the
// orchestration of the kernels behind each entry point.  No allocation, no sync: every
// call only enqueues kernels on the caller's stream.
#include "../../include/nrms_b200.h"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>

#include "attention.cuh"
#include "attention_mma.cuh"
#include "attention_hp.cuh"
#include "attention_hpn.cuh"
#include "attention_hpl.cuh"
#include "common.cuh"
#include "gemm_simt.cuh"
#include "gather.cuh"
#include "gemm_img.cuh"
#include "masked_user.cuh"
#include "metrics.cuh"
#include "pooling.cuh"
#include "profiler.cuh"
#include "train_kernels.cuh"

using namespace nrms;

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define NRMS_CHECK_CUDA(expr)                                                              \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess)                                                             \
            return fail(NRMS_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                        __FILE__, __LINE__);                                               \
    } while (0)

#define NRMS_REQUIRE_PTR(p)                                                    \
    do {                                                                       \
        if ((p) == nullptr) return fail(NRMS_ERR_NULL, "%s is NULL", #p);      \
        if ((reinterpret_cast<uintptr_t>(p) & 15u) != 0)                       \
            return fail(NRMS_ERR_ALIGN, "%s is not 16-byte aligned", #p);      \
    } while (0)

inline int grid_for(long long work_items, int threads, int max_waves = 8) {
    long long b = ceil_div64(work_items, threads);
    long long cap = (long long)kNumSMs * max_waves;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

// ---- flat parameter block -------------------------------------------------------------
struct ParamView {
    const float* Wqkv;  // [3D, D]
    const float* bqkv;  // [3D]
    const float* Wa;    // [Q, D]
    const float* ba;    // [Q]
    const float* qv;    // [Q]
};
struct GradView {
    float* Wqkv;
    float* bqkv;
    float* Wa;
    float* ba;
    float* qv;
};
template <typename V, typename P>
V param_view(P base, int D, int Q) {
    V v;
    v.Wqkv = base;
    v.bqkv = base + 3ll * D * D;
    v.Wa = v.bqkv + 3ll * D;
    v.ba = v.Wa + (long long)Q * D;
    v.qv = v.ba + Q;
    return v;
}

// ---- saved / scratch blob layouts -------------------------------------------------------
// gemm_mode 1 keeps every GEMM operand as a split-bf16 image (gemm_img.cuh); gemm_mode 0 keeps
// them as fp32 matrices for the CUDA-core GEMM.  Image sizes are fixed by the tile shapes of
// the kernels (rows / chunks an operand tile may touch), not only by the matrix dims.
constexpr int kXChunks = 5;      // D <= 320: K-major k-chunks and the 5 blocks of an N=320 tile
constexpr int kPreChunks = 4;    // Q <= 208
constexpr int kQkvChunks = 16;   // 3D <= 960 (+ the second block of the last 128-row tile)
constexpr int kWqkvRows = 1024, kWaRows = 256;

inline int mask_bytes_for(int D) { return (int)align_up(ceil_div(D, 8), 4); }

// Head-padded Q|K|V ("HP", gemm_img.cuh: hp_unpad; attention_hp.cuh, attention_hpn.cuh, attention_hpl.cuh):
// tensor-core GEMM modes, sequences of at most 256 tokens, even head dim <= 32, 3*32*h <= 960 projection
// columns.
inline bool use_hp(const nrms_encoder_dims& d) {
    const int dk = d.d_model / d.n_heads;
    return d.gemm_mode >= 1 && d.seq_len <= 256 && dk % 2 == 0 && dk <= 32 && 96 * d.n_heads <= 960;
}
inline int hp_cols(const nrms_encoder_dims& d) { return 96 * d.n_heads; }
// rows per head block: 32 / 64 (one- and several-warp kernels), a multiple of 16 beyond (key-tiled kernels)
inline int hp_rows(const nrms_encoder_dims& d) {
    return d.seq_len <= 32 ? 32 : d.seq_len <= 64 ? 64 : (int)align_up(d.seq_len, 16);
}

// CTA pairs (cta_group::2) for the forward Q|K|V projection (gemm_img.cuh: PAIR): opt-in with NRMS_PAIRS=1.
// Correct (tests/test_gpu_gemm.py) but MEASURED SLOWER than one CTA per SM on this path (cfg2: 0.229 ms vs
// 0.198 ms, cfg3: 1.26 vs 1.14 ms; profiles/r02_pairs_ab.txt): the projection's main loop already runs at
// the 3-term tensor roofline, so halving the weight-tile traffic buys nothing and the pair's lock-step costs.
inline bool use_pairs() {
    static const bool on = getenv("NRMS_PAIRS") && atoi(getenv("NRMS_PAIRS")) == 1;
    return on;
}

// sequences of at least this many tokens take the key-tiled kernels of attention_hpl.cuh (experiment knobs:
// NRMS_HPL_MIN_FWD / NRMS_HPL_MIN_BWD; the defaults are the measured cross-over points)
inline int hpl_min(bool fwd) {
    static const int f = getenv("NRMS_HPL_MIN_FWD") ? atoi(getenv("NRMS_HPL_MIN_FWD")) : 33;
    static const int b = getenv("NRMS_HPL_MIN_BWD") ? atoi(getenv("NRMS_HPL_MIN_BWD")) : 65;
    return fwd ? f : b;
}

// 48-row tiles for sequences of 33..48 tokens (experiment knobs: NRMS_HPN48=0 -> 64-row backward tiles and the
// key-tiled forward as before; NRMS_HPN48_FWD=0 -> only the forward as before).  Measured on cfg5 (48-token
// titles, bf16 products): backward 3.34 -> 2.31 ms, forward 2.22 -> 1.76 ms, step 12.6 -> 11.2 ms.
inline bool hpn48() {
    static const bool on = !(getenv("NRMS_HPN48") && atoi(getenv("NRMS_HPN48")) == 0);
    return on;
}

inline bool hpn48_fwd() {
    static const bool on = !(getenv("NRMS_HPN48_FWD") && atoi(getenv("NRMS_HPN48_FWD")) == 0);
    return on && hpn48();
}

// Second-generation latency-oriented kernels (pooling.cuh: pool_fwd2 / pool_bwd2 / reduce_rows_w).  Measured A/B on
// one box (scripts/ab_v2.sh, profiles/r02_v2_ab.txt; cfg2, per-launch times of the serialised pass):
//   * row reducer: 12.8 + 11.5 -> 7.8 + 10.7 us (title encoder), 9.0 -> 6.6 us (user encoder): on by default;
//   * pooling: the user encoder's launches (64 CTAs: a single CTA's serial chain is the whole cost) 11.3 -> 9.6 us
//     forward, 17.0 -> 14.8 us backward; the title encoder's 3,520-CTA launches are throughput-bound (backward
//     4.2 TB/s) and the 80-register v2 kernels run 3 CTAs per SM instead of 8: backward 78 -> 91 us, forward
//     unchanged — so v2 pooling is used for launches that cannot fill the GPU (kPool2MaxSeq);
//   * four (instead of two) gather items in flight per thread: 64 -> 72 us; not kept.
// NRMS_V2 is a bit mask for A/B runs: 1 = pooling forward, 2 = pooling backward (both: also for large launches),
// 4 = row reducer; unset = the measured default described above.
constexpr int kV2PoolFwd = 1, kV2PoolBwd = 2, kV2Reduce = 4;
constexpr int kPool2MaxSeq = 1024;
inline int v2_mask() {
    static const int mask = getenv("NRMS_V2") ? atoi(getenv("NRMS_V2")) : -1;
    return mask;
}
inline bool v2_reduce() { return v2_mask() < 0 || (v2_mask() & kV2Reduce) != 0; }
inline bool v2_pool(int bit, int n_seq) { return v2_mask() < 0 ? n_seq <= kPool2MaxSeq : (v2_mask() & bit) != 0; }

struct Saved {
    float* qkv;      // [M, 3D] fp32; HP: bf16 planes hi [M, NP] then lo [M, NP]
    float* lse;      // [M, h]
    float* ctx;      // [M, D]   (post-dropout)
    float* t;        // [M, Q]
    float* score;    // [M]      a_l = t_l . q
    float* w;        // [n_seq, L]
    uint8_t* xmask;  // [M, mask_bytes] embedding-dropout keep bits
    uint8_t* cmask;  // [M, mask_bytes] context-dropout keep bits
    float* x_f32;    // mode 0, news encoder: gathered + dropped rows [M, D]
    ig::Img x_img, ctx_img, wqkv_img, wa_img;   // mode 1
    int64_t bytes;
};
Saved saved_layout(void* blob, const nrms_encoder_dims& d) {
    const int64_t M = (int64_t)d.n_seq * d.seq_len;
    const int D = d.d_model, Q = d.d_query;
    char* p = reinterpret_cast<char*>(blob);
    int64_t off = 0;
    auto take_bytes = [&](int64_t n) {
        char* r = p + off;
        off += align_up(n, 1024);
        return r;
    };
    auto take = [&](int64_t nfloat) { return reinterpret_cast<float*>(take_bytes(nfloat * 4)); };
    Saved s{};
    s.qkv = use_hp(d) ? take((int64_t)d.n_seq * hp_rows(d) * hp_cols(d)) : take(M * 3 * D);   // HP: 32/64-row blocks
    s.lse = take(M * d.n_heads);
    s.ctx = take(M * D);
    s.t = take(M * Q);
    s.score = take(M);
    s.w = take(M);
    const int mb = mask_bytes_for(D);
    s.xmask = reinterpret_cast<uint8_t*>(take_bytes(M * mb));
    s.cmask = reinterpret_cast<uint8_t*>(take_bytes(M * mb));
    if (d.gemm_mode >= 1) {
        s.x_img = ig::img_view(take_bytes(ig::img_bytes(M, kXChunks)), M, kXChunks);
        s.ctx_img = ig::img_view(take_bytes(ig::img_bytes(M, kXChunks)), M, kXChunks);
        // gemm_mode 2 (plain bf16 products): activations and gradients are SINGLE-plane images — no kernel of
        // that mode reads a lo plane, so none is written (the blob keeps the two-plane size)
        if (d.gemm_mode == 2 && blob != nullptr) s.x_img.lo = s.ctx_img.lo = nullptr;
        s.wqkv_img = ig::img_view(take_bytes(ig::img_bytes(kWqkvRows, kXChunks)), kWqkvRows, kXChunks);
        s.wa_img = ig::img_view(take_bytes(ig::img_bytes(kWaRows, kXChunks)), kWaRows, kXChunks);
    } else {
        s.x_f32 = take(M * D);
    }
    s.bytes = off;
    return s;
}

constexpr int kReduceSlices = 128;

// k-splits of a weight-gradient GEMM: fill the SMs, never more splits than 64-row k chunks
int wgrad_splits_tc(int m_tiles, int k_chunks) {
    int s = kNumSMs / m_tiles;
    if (s > k_chunks) s = k_chunks;
    if (s < 1) s = 1;
    // the kernel gives every split ceil(k_chunks/s) chunks: drop the splits that would be empty
    return ceil_div(k_chunks, ceil_div(k_chunks, s));
}
int wgrad_splits_simt(int out_rows, int out_cols, long long k_rows) {
    const int tiles = ceil_div(out_rows, GBM) * ceil_div(out_cols, GBN);
    int s = ceil_div(4 * kNumSMs, tiles);
    if (s > 32) s = 32;
    const long long max_by_k = ceil_div64(k_rows, 4 * GBK);
    if (s > max_by_k) s = (int)max_by_k;
    return s < 1 ? 1 : s;
}

struct Scratch {
    float* d_ctx;      // [M, D]
    float* d_pre;      // mode 0: [M, Q]
    float* d_qkv;      // mode 0: [M, 3D]
    ig::Img d_pre_img, d_qkv_img;   // mode 1
    float* part_q;     // [n_seq, 2Q]
    float* part_b;     // [n_seq, 3D]
    float* wpart;      // [splits, out*in] weight-gradient partials
    float* red_tmp;    // [kReduceSlices, max(3D, 2Q)]
    int64_t bytes;
};
Scratch scratch_layout(void* blob, const nrms_encoder_dims& d) {
    const int64_t M = (int64_t)d.n_seq * d.seq_len;
    const int64_t D = d.d_model, Q = d.d_query;
    char* p = reinterpret_cast<char*>(blob);
    int64_t off = 0;
    auto take_bytes = [&](int64_t n) {
        char* r = p + off;
        off += align_up(n, 1024);
        return r;
    };
    auto take = [&](int64_t nfloat) { return reinterpret_cast<float*>(take_bytes(nfloat * 4)); };
    Scratch s{};
    s.d_ctx = take(M * D);
    int64_t wpart;
    if (d.gemm_mode >= 1) {
        s.d_pre_img = ig::img_view(take_bytes(ig::img_bytes(M, kPreChunks)), M, kPreChunks);
        s.d_qkv_img = ig::img_view(take_bytes(ig::img_bytes(M, kQkvChunks)), M, kQkvChunks);
        if (d.gemm_mode == 2 && blob != nullptr) s.d_pre_img.lo = s.d_qkv_img.lo = nullptr;
        const int kch = ig::img_rows_pad(M) / 64;
        const int64_t nq = use_hp(d) ? hp_cols(d) : 3 * D;   // output rows of dW_qkv (HP: padded order)
        const int64_t a = (int64_t)wgrad_splits_tc(ceil_div((int)nq, 128), kch) * nq * (D + 4);
        const int64_t b = (int64_t)wgrad_splits_tc(ceil_div((int)Q, 128), kch) * Q * (D + 4);
        wpart = a > b ? a : b;
    } else {
        s.d_pre = take(M * Q);
        s.d_qkv = take(M * 3 * D);
        wpart = 32 * 3 * D * (D > Q ? D : Q);
    }
    s.part_q = take((int64_t)d.n_seq * 2 * Q);
    s.part_b = take((int64_t)d.n_seq * 3 * D);
    s.wpart = take(wpart);
    s.red_tmp = take((int64_t)kReduceSlices * (3 * D > 2 * Q ? 3 * D : 2 * Q));
    s.bytes = off;
    return s;
}

int check_dims(const nrms_encoder_dims* d, bool news) {
    if (!d) return fail(NRMS_ERR_NULL, "dims is NULL");
    if (d->n_seq < 1 || d->seq_len < 1)
        return fail(NRMS_ERR_BAD_SHAPE, "n_seq=%d seq_len=%d must be >= 1", d->n_seq, d->seq_len);
    if (d->seq_len > 256) return fail(NRMS_ERR_BAD_SHAPE, "seq_len=%d > 256 unsupported", d->seq_len);
    if (d->d_model < 4 || d->d_model % 4 || d->d_model > 384)
        return fail(NRMS_ERR_BAD_SHAPE, "d_model=%d must be a multiple of 4 in [4,384]", d->d_model);
    if (d->d_query < 4 || d->d_query % 4)
        return fail(NRMS_ERR_BAD_SHAPE, "d_query=%d must be a positive multiple of 4", d->d_query);
    if (d->n_heads < 1 || d->d_model % d->n_heads)
        return fail(NRMS_ERR_BAD_SHAPE, "d_model=%d not divisible by n_heads=%d", d->d_model,
                    d->n_heads);
    if (d->d_model / d->n_heads > kDkPad)
        return fail(NRMS_ERR_BAD_SHAPE, "head dim %d > %d unsupported", d->d_model / d->n_heads,
                    kDkPad);
    if (news && d->vocab < 1) return fail(NRMS_ERR_BAD_SHAPE, "vocab=%d must be >= 1", d->vocab);
    if (d->dropout_p < 0.f || d->dropout_p >= 1.f)
        return fail(NRMS_ERR_BAD_SHAPE, "dropout_p=%f outside [0,1)", (double)d->dropout_p);
    if ((int64_t)d->n_seq * d->seq_len > 0x7fffffffll / 4)
        return fail(NRMS_ERR_BAD_SHAPE, "n_seq*seq_len too large");
    if (d->gemm_mode < 0 || d->gemm_mode > 2)
        return fail(NRMS_ERR_BAD_SHAPE, "gemm_mode=%d unknown", d->gemm_mode);
    if (d->gemm_mode >= 1) {
        // tile shapes of the tcgen05 path (gemm_img.cuh): N tiles of 240 / 208 / 320 columns
        if (d->d_model > 316 || d->d_query > 208)
            return fail(NRMS_ERR_BAD_SHAPE, "gemm_mode 1 supports d_model <= 316 and d_query <= 208 "
                        "(got %d, %d); use gemm_mode 0", d->d_model, d->d_query);
        if ((d->d_model / d->n_heads) % 2)
            return fail(NRMS_ERR_BAD_SHAPE, "gemm_mode 1 needs an even head dim (got %d); use gemm_mode 0",
                        d->d_model / d->n_heads);
    }
    return NRMS_OK;
}

// deterministic column sums: out[n] = scale * sum_r in[r, n]
int reduce_rows(const float* in, float* out, long long R, long long n, long long ld, float scale,
                float* tmp, cudaStream_t s) {
    const int threads = 256;
    if (v2_reduce()) {
        // rows over the warps of a CTA (reduce_rows_w_kernel); two levels when there are enough rows to fill SMs
        const dim3 blk(32, kRedWarps);
        const unsigned gx = (unsigned)ceil_div64(n, 32);
        if (R <= 4 * kRedWarps || tmp == nullptr) {
            NRMS_LAUNCH("reduce_rows", s, reduce_rows_w_kernel<<<dim3(gx, 1), blk, 0, s>>>(in, out, R, n, ld, R, scale));
        } else {
            const long long per = ceil_div64(R, kReduceSlices);
            const int slices = (int)ceil_div64(R, per);
            NRMS_LAUNCH("reduce_rows_sliced", s, reduce_rows_w_kernel<<<dim3(gx, slices), blk, 0, s>>>(in, tmp, R, n, ld, per, 1.f));
            NRMS_LAUNCH("reduce_rows", s, reduce_rows_w_kernel<<<dim3(gx, 1), blk, 0, s>>>(tmp, out, slices, n, n, slices, scale));
        }
        NRMS_CHECK_CUDA(cudaGetLastError());
        return NRMS_OK;
    }
    if (R <= 256 || tmp == nullptr) {
        NRMS_LAUNCH("reduce_rows", s, reduce_rows_kernel<<<(unsigned)ceil_div64(n, threads), threads, 0, s>>>(in, out, R, n, ld,
                                                                               scale, 0));
    } else {
        const long long per = ceil_div64(R, kReduceSlices);
        const int slices = (int)ceil_div64(R, per);
        NRMS_LAUNCH("reduce_rows_sliced", s, reduce_rows_sliced_kernel<<<dim3((unsigned)ceil_div64(n, threads), slices), threads, 0, s>>>(
            in, tmp, R, n, ld, per));
        NRMS_LAUNCH("reduce_rows", s, reduce_rows_kernel<<<(unsigned)ceil_div64(n, threads), threads, 0, s>>>(tmp, out, slices, n,
                                                                               n, scale, 0));
    }
    NRMS_CHECK_CUDA(cudaGetLastError());
    return NRMS_OK;
}

// ---- attention launch configuration ---------------------------------------------------------
struct AttnCfg {
    int hpb, threads, unit;
    size_t smem;
};
AttnCfg attn_fwd_cfg(int L, int n_heads, int n_seq) {
    AttnCfg c;
    c.unit = L <= 32 ? 16 : 32;                       // lanes per (head, row-block) task; 2 rows per lane
    const int rb = ceil_div(L, 2 * c.unit);
    const size_t budget = 110 * 1024;                 // 2 CTAs per SM (registers allow no more)
    int hpb = n_heads;
    while (hpb > 1 && (attn_fwd_smem_bytes(L, hpb) > budget || hpb * rb * c.unit > 256)) --hpb;
    (void)n_seq;   // (splitting the head groups further for few sequences measured slower)
    hpb = ceil_div(n_heads, ceil_div(n_heads, hpb));  // balance the head groups
    c.hpb = hpb;
    int threads = hpb * rb * c.unit;
    if (threads > 256) threads = 256;
    c.threads = (int)align_up(threads, 32);
    c.smem = attn_fwd_smem_bytes(L, hpb);
    return c;
}
AttnCfg attn_bwd_cfg(int L, int n_heads, int n_seq) {
    AttnCfg c;
    c.unit = 32;
    const int rc = ceil_div(L, 32);
    const size_t budget = 110 * 1024;                 // 2 CTAs per SM
    int hpb = n_heads;
    while (hpb > 1 && (hpb * rc > 8 || attn_bwd_smem_bytes(L, hpb) > budget)) --hpb;
    (void)n_seq;
    hpb = ceil_div(n_heads, ceil_div(n_heads, hpb));  // balance the head groups
    c.hpb = hpb;
    c.threads = hpb * rc * 32;
    c.smem = attn_bwd_smem_bytes(L, hpb);
    return c;
}

template <typename K>
int set_smem(K kernel, size_t smem) {
    NRMS_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    return NRMS_OK;
}

ig::IgArgs ig_args(const ig::Img& A, const ig::Img& B, float* C, int ldc, int M, int N) {
    ig::IgArgs g{};
    g.A = A; g.B = B; g.C = C; g.ldc = ldc; g.M = M; g.N = N;
    g.splits = 1;
    g.terms = 3;
    g.mask_scale = 1.f;
    return g;
}

int encoder_fwd(const nrms_encoder_dims& d, const int64_t* ids, const float* x_or_table,
                const float* params, float* out, void* saved_blob, int64_t saved_bytes,
                bool news, cudaStream_t s) {
    const ProfileScope prof_scope(news ? "" : "@user");
    const Saved sv = saved_layout(saved_blob, d);
    if (saved_bytes < sv.bytes)
        return fail(NRMS_ERR_WORKSPACE, "saved blob %lld < %lld bytes", (long long)saved_bytes,
                    (long long)sv.bytes);
    const int D = d.d_model, Q = d.d_query, L = d.seq_len, h = d.n_heads, dk = D / h;
    const int M = d.n_seq * L;
    const bool tcm = d.gemm_mode >= 1;
    const bool hp = use_hp(d);
    const int NP = hp_cols(d);
    const int terms = d.gemm_mode == 2 ? 1 : 3;   // mode 2: plain bf16 tensor-core products
    const ParamView pv = param_view<ParamView>(params, D, Q);
    const Dropout drop = make_dropout(d.dropout_p, d.seed);
    const int mb = mask_bytes_for(D);
    int rc;

    // 1. embedding gather + embedding dropout (nrms_v0.py:166) -> GEMM operand.  The user
    //    encoder's input is already a matrix: identity gather into an image (mode 1 only).
    const float* x_f32 = x_or_table;
    if (!tcm && !news && ids != nullptr)
        return fail(NRMS_ERR_BAD_SHAPE, "the gathering user encoder needs a tensor-core GEMM mode (gemm_mode >= 1)");
    if (news || tcm) {
        GatherArgs g{};
        // ids == nullptr: identity (the user encoder over a dense [n_seq*L, D] input); the user encoder may also
        // gather its rows from a vector table by id (cached-vector scoring: nrms_user_encoder_fwd_gather)
        g.table = x_or_table; g.ids = ids; g.M = M; g.vocab = ids ? d.vocab : M; g.D = D;
        g.x_f32 = tcm ? nullptr : sv.x_f32;
        if (tcm) g.x_img = sv.x_img;
        g.mask = sv.xmask; g.mask_bytes = mb;
        g.drop = news ? drop : make_dropout(0.f, 0);
        const long long items = tcm ? (long long)sv.x_img.rows_pad * sv.x_img.chunks * 8 : (long long)M * ceil_div(D, 8);
        NRMS_LAUNCH("gather", s, gather_rows_img_kernel<2><<<grid_for(items, 256, 16), 256, 0, s>>>(g));
        NRMS_CHECK_CUDA(cudaGetLastError());
        x_f32 = sv.x_f32;
    }
    // 2. Q|K|V projections (nrms_v0.py:53-58)
    if (tcm && hp) {
        // head-padded projection: weight image rows in padded order, output = split-bf16 planes
        NRMS_CHECK_CUDA(ig::img_pack2(pv.Wqkv, 3 * D, D, D, sv.wqkv_img, D, dk, pv.Wa, Q, D, D, sv.wa_img, s));
        ig::IgArgs g = ig_args(sv.x_img, sv.wqkv_img, nullptr, NP, M, NP);
        g.Chi = reinterpret_cast<uint16_t*>(sv.qkv);
        g.Clo = terms == 3 ? g.Chi + (long long)d.n_seq * hp_rows(d) * NP : nullptr;   // plain bf16: hi plane only
        g.hp_D = D; g.hp_dk = dk; g.seq_len = L; g.hp_rows = hp_rows(d);
        g.terms = terms;
        g.bias = pv.bqkv;
        g.m_tiles = sv.x_img.rows_pad / 128; g.n_tiles = ceil_div(NP, 256);
        g.k_steps = ceil_div(D, 16); g.k_chunks = ceil_div(g.k_steps, 4);
        if (use_pairs())
            NRMS_CHECK_CUDA((ig::ig_launch<false, false, 256, ig::EPI_BIAS_SPLIT, true>(g, s, "gemm_fwd_qkv")));
        else
            NRMS_CHECK_CUDA((terms == 3 ? ig::ig_launch<false, false, 256, ig::EPI_BIAS_SPLIT>(g, s, "gemm_fwd_qkv")
                                        : ig::ig_launch<false, false, 256, ig::EPI_BIAS_SPLIT_HI>(g, s, "gemm_fwd_qkv")));
    } else if (tcm) {
        NRMS_CHECK_CUDA(ig::img_pack2(pv.Wqkv, 3 * D, D, D, sv.wqkv_img, 0, 0, pv.Wa, Q, D, D, sv.wa_img, s));
        ig::IgArgs g = ig_args(sv.x_img, sv.wqkv_img, sv.qkv, 3 * D, M, 3 * D);
        g.terms = terms;
        g.bias = pv.bqkv;
        g.m_tiles = sv.x_img.rows_pad / 128; g.n_tiles = ceil_div(3 * D, 240);
        g.k_steps = ceil_div(D, 16); g.k_chunks = ceil_div(g.k_steps, 4);
        NRMS_CHECK_CUDA((ig::ig_launch<false, false, 240, ig::EPI_BIAS>(g, s, "gemm_fwd_qkv")));
    } else {
        GemmArgs g{};
        g.A = x_f32; g.B = pv.Wqkv; g.C = sv.qkv; g.bias = pv.bqkv;
        g.M = M; g.N = 3 * D; g.K = D; g.lda = D; g.ldb = D; g.ldc = 3 * D; g.k_chunk = D;
        NRMS_CHECK_CUDA(launch_gemm_simt(g, true, true, 1, s, "gemm_fwd_qkv"));
    }
    // 3. per-head attention (+ context dropout for the news encoder)
    {
        AttnArgs a{};
        a.qkv = sv.qkv; a.ctx = sv.ctx; a.lse = sv.lse; a.cmask = sv.cmask; a.mask_bytes = mb;
        if (tcm) a.ctx_img = sv.ctx_img;
        a.M = M; a.L = L; a.D = D; a.n_heads = h; a.dk = dk;
        a.scale = 1.f / sqrtf((float)dk);
        a.drop = news ? drop : make_dropout(0.f, 0);
        if (hp) {
            // one independent warp per (sequence, head) over the head-padded planes (attention_hp.cuh)
            a.ctx = nullptr;    // the context lives in its image only: the pooling kernels read hi + lo
            a.qkv = nullptr;
            a.qkv_hi = reinterpret_cast<const uint16_t*>(sv.qkv);
            a.qkv_lo = a.qkv_hi + (long long)d.n_seq * hp_rows(d) * NP;
            const long long items = (long long)d.n_seq * h;
            if (L > 32 && L <= 48 && hpn48_fwd()) {
                // three warps per (sequence, head) over the first 48 rows of the 64-row blocks (attention_hpn.cuh)
                using C = HpN<48>;
                const size_t smem = (size_t)C::ITEMS_FWD * C::ITEM_FWD;
                const unsigned grid = (unsigned)std::min<long long>(ceil_div64(items, C::ITEMS_FWD), (terms == 3 ? 3 : 4) * kNumSMs);
                const int threads = C::ITEMS_FWD * C::NW * 32;
                if (terms == 3) {
                    if ((rc = set_smem(attn_hpn_fwd_kernel<3, 48>, smem))) return rc;
                    NRMS_LAUNCH("attn_fwd", s, (attn_hpn_fwd_kernel<3, 48><<<grid, threads, smem, s>>>(a, items)));
                } else {
                    if ((rc = set_smem(attn_hpn_fwd_kernel<1, 48>, smem))) return rc;
                    NRMS_LAUNCH("attn_fwd", s, (attn_hpn_fwd_kernel<1, 48><<<grid, threads, smem, s>>>(a, items)));
                }
            } else if (L >= hpl_min(true)) {
                // one CTA per (sequence, head), a warp per 16 rows, keys in tiles of 64 with an online softmax
                // (attention_hpl.cuh)
                const size_t smem = attn_hpl_fwd_smem_bytes(L);
                const unsigned grid = (unsigned)std::min<long long>(items, 8ll * kNumSMs);
                const int threads = hpl_warps(L) * 32;
                if (terms == 3) {
                    if ((rc = set_smem(attn_hpl_fwd_kernel<3>, smem))) return rc;
                    NRMS_LAUNCH("attn_fwd", s, (attn_hpl_fwd_kernel<3><<<grid, threads, smem, s>>>(a, items, hp_rows(d))));
                } else {
                    if ((rc = set_smem(attn_hpl_fwd_kernel<1>, smem))) return rc;
                    NRMS_LAUNCH("attn_fwd", s, (attn_hpl_fwd_kernel<1><<<grid, threads, smem, s>>>(a, items, hp_rows(d))));
                }
            } else if (L > 32) {
                // four warps per (sequence, head) over 64-row blocks (attention_hpn.cuh)
                using C = HpN<64>;
                const size_t smem = (size_t)C::ITEMS_FWD * C::ITEM_FWD;
                const unsigned grid = (unsigned)std::min<long long>(ceil_div64(items, C::ITEMS_FWD), 4 * kNumSMs);
                const int threads = C::ITEMS_FWD * C::NW * 32;
                if (terms == 3) {
                    if ((rc = set_smem(attn_hpn_fwd_kernel<3, 64>, smem))) return rc;
                    NRMS_LAUNCH("attn_fwd", s, (attn_hpn_fwd_kernel<3, 64><<<grid, threads, smem, s>>>(a, items)));
                } else {
                    if ((rc = set_smem(attn_hpn_fwd_kernel<1, 64>, smem))) return rc;
                    NRMS_LAUNCH("attn_fwd", s, (attn_hpn_fwd_kernel<1, 64><<<grid, threads, smem, s>>>(a, items)));
                }
            } else {
            const size_t smem = attn_hp_fwd_smem_bytes();
            const unsigned grid = (unsigned)std::min<long long>(ceil_div64(items, kHpFwdWarps), 2 * kNumSMs);   // persistent warps
            if (terms == 3) {
                if ((rc = set_smem(attn_hp_fwd_kernel<3>, smem))) return rc;
                NRMS_LAUNCH("attn_fwd", s, (attn_hp_fwd_kernel<3><<<grid, kHpFwdWarps * 32, smem, s>>>(a, items)));
            } else {
                if ((rc = set_smem(attn_hp_fwd_kernel<1>, smem))) return rc;
                NRMS_LAUNCH("attn_fwd", s, (attn_hp_fwd_kernel<1><<<grid, kHpFwdWarps * 32, smem, s>>>(a, items)));
            }
            }
        } else if (L <= kTile && dk % 2 == 0) {
            // one independent warp per (sequence, head), products on mma.sync (attention_mma.cuh)
            const long long items = (long long)d.n_seq * h;
            const size_t smem = attn_mma_fwd_smem_bytes();
            const unsigned grid = (unsigned)ceil_div64(items, kMmaWarpsFwd);
            if (terms == 3) {
                if ((rc = set_smem(attn_mma_fwd_kernel<3>, smem))) return rc;
                NRMS_LAUNCH("attn_fwd", s, (attn_mma_fwd_kernel<3><<<grid, kMmaWarpsFwd * 32, smem, s>>>(a, items)));
            } else {
                if ((rc = set_smem(attn_mma_fwd_kernel<1>, smem))) return rc;
                NRMS_LAUNCH("attn_fwd", s, (attn_mma_fwd_kernel<1><<<grid, kMmaWarpsFwd * 32, smem, s>>>(a, items)));
            }
        } else {
            const AttnCfg c = attn_fwd_cfg(L, h, d.n_seq);
            a.hpb = c.hpb;
            const bool vec2 = (dk % 2 == 0);
            const dim3 grid(d.n_seq, ceil_div(h, c.hpb));
            if (c.unit == 16) {
                if (vec2) {
                    if ((rc = set_smem(attn_fwd_kernel<16, true>, c.smem))) return rc;
                    NRMS_LAUNCH("attn_fwd", s, (attn_fwd_kernel<16, true><<<grid, c.threads, c.smem, s>>>(a)));
                } else {
                    if ((rc = set_smem(attn_fwd_kernel<16, false>, c.smem))) return rc;
                    NRMS_LAUNCH("attn_fwd", s, (attn_fwd_kernel<16, false><<<grid, c.threads, c.smem, s>>>(a)));
                }
            } else {
                if (vec2) {
                    if ((rc = set_smem(attn_fwd_kernel<32, true>, c.smem))) return rc;
                    NRMS_LAUNCH("attn_fwd", s, (attn_fwd_kernel<32, true><<<grid, c.threads, c.smem, s>>>(a)));
                } else {
                    if ((rc = set_smem(attn_fwd_kernel<32, false>, c.smem))) return rc;
                    NRMS_LAUNCH("attn_fwd", s, (attn_fwd_kernel<32, false><<<grid, c.threads, c.smem, s>>>(a)));
                }
            }
        }
        NRMS_CHECK_CUDA(cudaGetLastError());
    }
    // 4. additive-attention projection t = tanh(ctx W_a^T + b_a) (nrms_v0.py:108); the tcgen05
    //    epilogue also reduces a_l = t_l . q (nrms_v0.py:110)
    if (tcm) {
        ig::IgArgs g = ig_args(sv.ctx_img, sv.wa_img, sv.t, Q, M, Q);
        g.terms = terms;
        g.bias = pv.ba; g.qv = pv.qv; g.dot_out = sv.score;
        g.m_tiles = sv.ctx_img.rows_pad / 128; g.n_tiles = 1;
        g.k_steps = ceil_div(D, 16); g.k_chunks = ceil_div(g.k_steps, 4);
        NRMS_CHECK_CUDA((ig::ig_launch<false, false, 208, ig::EPI_TANH_DOT>(g, s, "gemm_fwd_additive")));
    } else {
        GemmArgs g{};
        g.A = sv.ctx; g.B = pv.Wa; g.C = sv.t; g.bias = pv.ba;
        g.M = M; g.N = Q; g.K = D; g.lda = D; g.ldb = D; g.ldc = Q; g.k_chunk = D; g.epilogue = 1;
        NRMS_CHECK_CUDA(launch_gemm_simt(g, true, true, 1, s, "gemm_fwd_additive"));
    }
    // 5. softmax over the sequence + weighted sum (nrms_v0.py:110-126)
    {
        PoolArgs p{};
        p.ctx = hp ? nullptr : sv.ctx; p.t = sv.t; p.q = pv.qv; p.w = sv.w; p.out = out;
        if (hp) p.ctx_img = sv.ctx_img;
        p.score = tcm ? sv.score : nullptr;
        p.M = M; p.L = L; p.D = D; p.Q = Q;
        const int pu = ceil_div(D, 8);
        const size_t pool_smem = (L + (hp ? 8 * pu * (256 / pu) : 0)) * sizeof(float);
        if (hp && tcm && v2_pool(kV2PoolFwd, d.n_seq))
            NRMS_LAUNCH("pool_fwd", s, pool_fwd2_kernel<<<d.n_seq, 256, pool_smem, s>>>(p));
        else
            NRMS_LAUNCH("pool_fwd", s, pool_fwd_kernel<<<d.n_seq, 256, pool_smem, s>>>(p));
        NRMS_CHECK_CUDA(cudaGetLastError());
    }
    return NRMS_OK;
}

int encoder_bwd(const nrms_encoder_dims& d, const int64_t* ids, const float* x_or_table,
                const float* params, const float* d_out, const void* saved_blob,
                int64_t saved_bytes, void* scratch_blob, int64_t scratch_bytes, float* d_params,
                float* d_x, bool news, int phases, cudaStream_t s) {
    (void)ids;
    const ProfileScope prof_scope(news ? "" : "@user");
    const Saved sv = saved_layout(const_cast<void*>(saved_blob), d);
    if (saved_bytes < sv.bytes)
        return fail(NRMS_ERR_WORKSPACE, "saved blob %lld < %lld bytes", (long long)saved_bytes,
                    (long long)sv.bytes);
    const Scratch sc = scratch_layout(scratch_blob, d);
    if (scratch_bytes < sc.bytes)
        return fail(NRMS_ERR_WORKSPACE, "scratch blob %lld < %lld bytes", (long long)scratch_bytes,
                    (long long)sc.bytes);
    const int D = d.d_model, Q = d.d_query, L = d.seq_len, h = d.n_heads, dk = D / h;
    const int M = d.n_seq * L;
    const bool tcm = d.gemm_mode >= 1;
    const bool hp = use_hp(d);
    const int NP = hp_cols(d);
    const int nq = hp ? NP : 3 * D;               // columns of the d_qkv image / rows of dW_qkv's partials
    const int terms = d.gemm_mode == 2 ? 1 : 3;
    const ParamView pv = param_view<ParamView>(params, D, Q);
    const GradView gv = param_view<GradView>(d_params, D, Q);
    const Dropout drop = make_dropout(news ? d.dropout_p : 0.f, d.seed);
    const int mb = mask_bytes_for(D);
    const float* x_f32 = news ? sv.x_f32 : x_or_table;   // mode 0 operand of dW_qkv
    int rc;

    const int tok_tiles = ig::img_rows_pad(M) / 128;
    const bool do_data = (phases & 1) != 0, do_params = (phases & 2) != 0;

    // ================= data-gradient path: d_out -> d_x (everything the caller's next step needs)
    if (do_data) {
        // 1. pooling backward: d_ctx (pool path), d_pre, partials of d_b_a and d_query
        {
            PoolArgs p{};
            p.ctx = hp ? nullptr : sv.ctx; p.t = sv.t; p.q = pv.qv; p.w = sv.w; p.d_out = d_out;
            if (hp) p.ctx_img = sv.ctx_img;
            p.d_part = sc.part_q;
            if (tcm) p.d_pre_img = sc.d_pre_img; else { p.d_pre = sc.d_pre; p.d_ctx = sc.d_ctx; }
            p.M = M; p.L = L; p.D = D; p.Q = Q;
            if (hp && tcm && Q % 8 == 0 && Q >= 8 && v2_pool(kV2PoolBwd, d.n_seq))
                NRMS_LAUNCH("pool_bwd", s, pool_bwd2_kernel<<<d.n_seq, 256, pool_bwd2_smem_floats(L, D, Q) * sizeof(float), s>>>(p));
            else
                NRMS_LAUNCH("pool_bwd", s, pool_bwd_kernel<<<d.n_seq, 256, 2 * L * sizeof(float), s>>>(p));
            NRMS_CHECK_CUDA(cudaGetLastError());
        }
        // 2. d_ctx = pool path + d_pre W_a
        if (tcm) {
            // d_ctx = w_l * d_out (pooling path, formed in the epilogue) + d_pre W_a
            ig::IgArgs g = ig_args(sc.d_pre_img, sv.wa_img, sc.d_ctx, D, M, D);
            g.terms = terms;
            g.m_tiles = tok_tiles; g.n_tiles = 1;
            g.k_steps = ceil_div(Q, 16); g.k_chunks = ceil_div(g.k_steps, 4);
            g.row_w = sv.w; g.seq_vec = d_out; g.seq_len = L;
            if (hp && drop.enabled()) {
                // the head-padded attention backward takes dO = d_ctx * keep/(1-p) ready-made: the context
                // dropout mask (nrms_v0.py:171-173) is applied here, where every row is already in registers
                g.mask_bits = reinterpret_cast<const uint32_t*>(sv.cmask);
                g.mask_words = mb / 4;
                g.mask_scale = drop.scale;
            }
            if (tok_tiles * 2 <= kNumSMs) {
                // few token tiles (the user encoder): 64-column tiles so that the grid fills the SMs
                g.n_tiles = ceil_div(D, 64);
                NRMS_CHECK_CUDA((ig::ig_launch<false, true, 64, ig::EPI_POOLADD>(g, s, "gemm_dgrad_additive")));
            } else
            NRMS_CHECK_CUDA((ig::ig_launch<false, true, 320, ig::EPI_POOLADD>(g, s, "gemm_dgrad_additive")));
        } else {
            GemmArgs g{};
            g.A = sc.d_pre; g.B = pv.Wa; g.C = sc.d_ctx;
            g.M = M; g.N = D; g.K = Q; g.lda = Q; g.ldb = D; g.ldc = D;
            g.k_chunk = Q; g.accumulate = 1;
            NRMS_CHECK_CUDA(launch_gemm_simt(g, true, false, 1, s, "gemm_dgrad_additive"));
        }
        // 4. attention backward -> d_qkv and bias partials
        {
            AttnArgs a{};
            a.qkv = sv.qkv; a.ctx = sv.ctx; a.lse = sv.lse; a.d_ctx = sc.d_ctx;
            a.cmask = sv.cmask; a.mask_bytes = mb;
            if (tcm) a.d_qkv_img = sc.d_qkv_img; else a.d_qkv = sc.d_qkv;
            a.d_bias_part = tcm ? nullptr : sc.part_b;   // tcgen05 path: bias gradient = column D of dW_qkv
            a.M = M; a.L = L; a.D = D; a.n_heads = h; a.dk = dk;
            a.scale = 1.f / sqrtf((float)dk);
            a.drop = drop;
            if (hp) {
                a.qkv = nullptr;
                a.qkv_hi = reinterpret_cast<const uint16_t*>(sv.qkv);
                a.qkv_lo = a.qkv_hi + (long long)d.n_seq * hp_rows(d) * NP;
                    a.cmask = nullptr;   // d_ctx arrives with the context-dropout mask applied (dgrad GEMM epilogue)
                const long long items = (long long)d.n_seq * h;
                // two (sequences of <= 32 tokens) or four warps per (sequence, head): attention_hpn.cuh;
                // beyond 64 tokens one CTA per item in two key-/query-tiled phases: attention_hpl.cuh
                if (L >= hpl_min(false)) {
                    a.ctx_img = sv.ctx_img;     // delta = dO . O comes from the saved context image
                    const size_t smem = attn_hpl_bwd_smem_bytes(L);
                    const unsigned grid = (unsigned)std::min<long long>(items, 8ll * kNumSMs);
                    const int threads = hpl_warps(L) * 32;
                    if (terms == 3) {
                        if ((rc = set_smem(attn_hpl_bwd_kernel<3>, smem))) return rc;
                        NRMS_LAUNCH("attn_bwd", s, (attn_hpl_bwd_kernel<3><<<grid, threads, smem, s>>>(a, items, hp_rows(d))));
                    } else {
                        if ((rc = set_smem(attn_hpl_bwd_kernel<1>, smem))) return rc;
                        NRMS_LAUNCH("attn_bwd", s, (attn_hpl_bwd_kernel<1><<<grid, threads, smem, s>>>(a, items, hp_rows(d))));
                    }
                } else if (L > 32 && L <= 48 && hpn48()) {
                    // 48-row tiles, three warps per item, four items per SM (52 KB each): a 48-token title on
                    // 64-row tiles wastes a quarter of the rows and 44 % of the S / P work
                    using C = HpN<48>;
                    const size_t smem = (size_t)C::ITEMS_BWD * C::item_bwd(terms);
                    const unsigned grid = (unsigned)std::min<long long>(ceil_div64(items, C::ITEMS_BWD), (terms == 3 ? 4 : 5) * kNumSMs);
                    const int threads = C::ITEMS_BWD * C::NW * 32;
                    if (terms == 3) {
                        if ((rc = set_smem(attn_hpn_bwd_kernel<3, 48>, smem))) return rc;
                        NRMS_LAUNCH("attn_bwd", s, (attn_hpn_bwd_kernel<3, 48><<<grid, threads, smem, s>>>(a, items)));
                    } else {
                        if ((rc = set_smem(attn_hpn_bwd_kernel<1, 48>, smem))) return rc;
                        NRMS_LAUNCH("attn_bwd", s, (attn_hpn_bwd_kernel<1, 48><<<grid, threads, smem, s>>>(a, items)));
                    }
                } else if (L > 32) {
                    using C = HpN<64>;
                    const size_t smem = (size_t)C::ITEMS_BWD * C::item_bwd(terms);
                    const unsigned grid = (unsigned)std::min<long long>(ceil_div64(items, C::ITEMS_BWD), (terms == 3 ? 2 : 3) * kNumSMs);
                    const int threads = C::ITEMS_BWD * C::NW * 32;
                    if (terms == 3) {
                        if ((rc = set_smem(attn_hpn_bwd_kernel<3, 64>, smem))) return rc;
                        NRMS_LAUNCH("attn_bwd", s, (attn_hpn_bwd_kernel<3, 64><<<grid, threads, smem, s>>>(a, items)));
                    } else {
                        if ((rc = set_smem(attn_hpn_bwd_kernel<1, 64>, smem))) return rc;
                        NRMS_LAUNCH("attn_bwd", s, (attn_hpn_bwd_kernel<1, 64><<<grid, threads, smem, s>>>(a, items)));
                    }
                } else {
                    using C = HpN<32>;
                    const size_t smem = (size_t)C::ITEMS_BWD * C::ITEM_BWD;
                    const unsigned grid = (unsigned)std::min<long long>(ceil_div64(items, C::ITEMS_BWD), 2 * kNumSMs);
                    const int threads = C::ITEMS_BWD * C::NW * 32;
                    if (terms == 3) {
                        if ((rc = set_smem(attn_hpn_bwd_kernel<3, 32>, smem))) return rc;
                        NRMS_LAUNCH("attn_bwd", s, (attn_hpn_bwd_kernel<3, 32><<<grid, threads, smem, s>>>(a, items)));
                    } else {
                        if ((rc = set_smem(attn_hpn_bwd_kernel<1, 32>, smem))) return rc;
                        NRMS_LAUNCH("attn_bwd", s, (attn_hpn_bwd_kernel<1, 32><<<grid, threads, smem, s>>>(a, items)));
                    }
                }
            } else if (L <= kTile && dk % 2 == 0) {
                const long long items = (long long)d.n_seq * h;
                const size_t smem = attn_mma_bwd_smem_bytes();
                const unsigned grid = (unsigned)ceil_div64(items, kMmaWarps);
                if (terms == 3) {
                    if ((rc = set_smem(attn_mma_bwd_kernel<3>, smem))) return rc;
                    NRMS_LAUNCH("attn_bwd", s, (attn_mma_bwd_kernel<3><<<grid, kMmaWarps * 32, smem, s>>>(a, items)));
                } else {
                    if ((rc = set_smem(attn_mma_bwd_kernel<1>, smem))) return rc;
                    NRMS_LAUNCH("attn_bwd", s, (attn_mma_bwd_kernel<1><<<grid, kMmaWarps * 32, smem, s>>>(a, items)));
                }
            } else {
                const AttnCfg c = attn_bwd_cfg(L, h, d.n_seq);
                a.hpb = c.hpb;
                const dim3 grid(d.n_seq, ceil_div(h, c.hpb));
                if (dk % 2 == 0) {
                    if ((rc = set_smem(attn_bwd_kernel<true>, c.smem))) return rc;
                    NRMS_LAUNCH("attn_bwd", s, (attn_bwd_kernel<true><<<grid, c.threads, c.smem, s>>>(a)));
                } else {
                    if ((rc = set_smem(attn_bwd_kernel<false>, c.smem))) return rc;
                    NRMS_LAUNCH("attn_bwd", s, (attn_bwd_kernel<false><<<grid, c.threads, c.smem, s>>>(a)));
                }
            }
            NRMS_CHECK_CUDA(cudaGetLastError());
        }
        // 6. d_x = d_qkv W_qkv (x the embedding dropout mask)
        if (tcm) {
            if (d_x) {
                ig::IgArgs g = ig_args(sc.d_qkv_img, sv.wqkv_img, d_x, D, M, D);
                g.terms = terms;
                g.m_tiles = tok_tiles; g.n_tiles = 1;
                g.k_steps = ceil_div(nq, 16); g.k_chunks = ceil_div(g.k_steps, 4);
                if (drop.enabled()) {
                    g.mask_bits = reinterpret_cast<const uint32_t*>(sv.xmask);
                    g.mask_words = mb / 4;
                    g.mask_scale = drop.scale;
                }
                if (tok_tiles * 2 <= kNumSMs) {
                    g.n_tiles = ceil_div(D, 64);
                    NRMS_CHECK_CUDA((ig::ig_launch<false, true, 64, ig::EPI_MASK>(g, s, "gemm_dgrad_qkv")));
                } else
                NRMS_CHECK_CUDA((ig::ig_launch<false, true, 320, ig::EPI_MASK>(g, s, "gemm_dgrad_qkv")));
            }
        } else {
            if (d_x) {
                GemmArgs g{};
                g.A = sc.d_qkv; g.B = pv.Wqkv; g.C = d_x;
                g.M = M; g.N = D; g.K = 3 * D; g.lda = 3 * D; g.ldb = D; g.ldc = D;
                g.k_chunk = 3 * D;
                if (drop.enabled()) {
                    g.mask = sv.xmask; g.mask_bytes = mb; g.mask_scale = drop.scale;
                }
                NRMS_CHECK_CUDA(launch_gemm_simt(g, true, false, 1, s, "gemm_dgrad_qkv"));
            }
        }
    }
    // ================= parameter gradients (off the data path: a data-parallel caller overlaps
    // them with the all-reduce of the embedding-table gradient)
    if (do_params) {
        // [d_b_a | d_query] are adjacent in the flat block, as in part_q; with the tcgen05 GEMMs d_b_a
        // comes out of the weight-gradient GEMM (ones column of the context image) and only d_query
        // is reduced here
        rc = tcm ? reduce_rows(sc.part_q + Q, gv.qv, d.n_seq, Q, 2 * Q, 1.f, sc.red_tmp, s)
                 : reduce_rows(sc.part_q, gv.ba, d.n_seq, 2 * Q, 2 * Q, 1.f, sc.red_tmp, s);
        if (rc) return rc;
        // 3. dW_a = d_pre^T ctx (split over the token rows)
        if (tcm) {
            // columns [0,D) = dW_a, column D = d_b_a (ones column of the context image)
            ig::IgArgs w = ig_args(sc.d_pre_img, sv.ctx_img, sc.wpart, D + 4, Q, D + 4);
            w.terms = terms;
            w.m_tiles = ceil_div(Q, 128); w.n_tiles = 1;
            w.k_chunks = tok_tiles * 2; w.k_steps = 4 * w.k_chunks;
            w.splits = wgrad_splits_tc(w.m_tiles, w.k_chunks);
            w.c_split_stride = (long long)Q * (D + 4);
            NRMS_CHECK_CUDA((ig::ig_launch<true, true, 320, ig::EPI_PARTIAL>(w, s, "gemm_wgrad_additive")));
            NRMS_LAUNCH("reduce_wgrad", s, reduce_wgrad_kernel<<<grid_for((long long)Q * (D / 4 + 1), 128), 128, 0, s>>>(
                sc.wpart, w.splits, Q, D + 4, D, gv.Wa, gv.ba, 0, 0));
            NRMS_CHECK_CUDA(cudaGetLastError());
        } else {
            const int splits = wgrad_splits_simt(Q, D, M);
            GemmArgs w{};
            w.A = sc.d_pre; w.B = sv.ctx; w.C = sc.wpart;
            w.M = Q; w.N = D; w.K = M; w.lda = Q; w.ldb = D; w.ldc = D;
            w.k_chunk = (int)align_up(ceil_div(M, splits), GBK);
            w.c_split_stride = (long long)Q * D;
            NRMS_CHECK_CUDA(launch_gemm_simt(w, false, false, ceil_div(M, w.k_chunk), s, "gemm_wgrad_additive"));
            rc = reduce_rows(sc.wpart, gv.Wa, ceil_div(M, w.k_chunk), (long long)Q * D,
                             (long long)Q * D, 1.f, nullptr, s);
            if (rc) return rc;
        }
        if (!tcm) {
            rc = reduce_rows(sc.part_b, gv.bqkv, d.n_seq, 3 * D, 3 * D, 1.f, sc.red_tmp, s);
            if (rc) return rc;
        }
        // 5. dW_qkv = d_qkv^T x
        if (tcm) {
            // columns [0,D) = dW_qkv, column D = d_b_qkv (ones column of the input image)
            // (HP: partial rows in head-padded order, mapped back by the reduction)
            ig::IgArgs w = ig_args(sc.d_qkv_img, sv.x_img, sc.wpart, D + 4, nq, D + 4);
            w.terms = terms;
            w.m_tiles = ceil_div(nq, 128); w.n_tiles = 1;
            w.k_chunks = tok_tiles * 2; w.k_steps = 4 * w.k_chunks;
            w.splits = wgrad_splits_tc(w.m_tiles, w.k_chunks);
            w.c_split_stride = (long long)nq * (D + 4);
            NRMS_CHECK_CUDA((ig::ig_launch<true, true, 320, ig::EPI_PARTIAL>(w, s, "gemm_wgrad_qkv")));
            NRMS_LAUNCH("reduce_wgrad", s, reduce_wgrad_kernel<<<grid_for((long long)nq * (D / 4 + 1), 128), 128, 0, s>>>(
                sc.wpart, w.splits, nq, D + 4, D, gv.Wqkv, gv.bqkv, hp ? D : 0, hp ? dk : 0));
            NRMS_CHECK_CUDA(cudaGetLastError());
        } else {
            const int splits = wgrad_splits_simt(3 * D, D, M);
            GemmArgs w{};
            w.A = sc.d_qkv; w.B = x_f32; w.C = sc.wpart;
            w.M = 3 * D; w.N = D; w.K = M; w.lda = 3 * D; w.ldb = D; w.ldc = D;
            w.k_chunk = (int)align_up(ceil_div(M, splits), GBK);
            w.c_split_stride = 3ll * D * D;
            NRMS_CHECK_CUDA(launch_gemm_simt(w, false, false, ceil_div(M, w.k_chunk), s, "gemm_wgrad_qkv"));
            rc = reduce_rows(sc.wpart, gv.Wqkv, ceil_div(M, w.k_chunk), 3ll * D * D, 3ll * D * D, 1.f,
                             nullptr, s);
            if (rc) return rc;
        }
    }
    return NRMS_OK;
}

}  // namespace

extern "C" {

int nrms_abi_version(void) { return NRMS_ABI_VERSION; }
const char* nrms_last_error(void) { return g_err; }

int64_t nrms_launch_count(void) { return Profiler::get().launches; }
void nrms_profile_enable(int on) { Profiler::get().on = on != 0; }
int nrms_profile_collect(char* h_buf, int64_t h_buf_bytes) {
    const std::string s = Profiler::get().collect();
    if (!h_buf || h_buf_bytes < 1) return fail(NRMS_ERR_NULL, "h_buf is NULL");
    const int64_t n = (int64_t)s.size() < h_buf_bytes - 1 ? (int64_t)s.size() : h_buf_bytes - 1;
    memcpy(h_buf, s.data(), (size_t)n);
    h_buf[n] = 0;
    return NRMS_OK;
}

int64_t nrms_encoder_param_count(int32_t D, int32_t Q) {
    return 3ll * D * D + 3ll * D + (int64_t)Q * D + 2ll * Q;
}
int64_t nrms_encoder_saved_bytes(const nrms_encoder_dims* d) {
    if (check_dims(d, false)) return -1;
    return saved_layout(nullptr, *d).bytes;
}
int64_t nrms_encoder_scratch_bytes(const nrms_encoder_dims* d) {
    if (check_dims(d, false)) return -1;
    return scratch_layout(nullptr, *d).bytes;
}

int nrms_news_encoder_fwd(const nrms_encoder_dims* d, const int64_t* ids, const float* table,
                          const float* params, float* out, void* saved, int64_t saved_bytes,
                          nrms_stream_t stream) {
    int rc = check_dims(d, true);
    if (rc) return rc;
    NRMS_REQUIRE_PTR(ids); NRMS_REQUIRE_PTR(table); NRMS_REQUIRE_PTR(params);
    NRMS_REQUIRE_PTR(out); NRMS_REQUIRE_PTR(saved);
    return encoder_fwd(*d, ids, table, params, out, saved, saved_bytes, true,
                       (cudaStream_t)stream);
}
int nrms_news_encoder_bwd(const nrms_encoder_dims* d, const int64_t* ids, const float* table,
                          const float* params, const float* d_out, const void* saved,
                          int64_t saved_bytes, void* scratch, int64_t scratch_bytes,
                          float* d_params, float* d_rows, nrms_stream_t stream) {
    int rc = check_dims(d, true);
    if (rc) return rc;
    NRMS_REQUIRE_PTR(ids); NRMS_REQUIRE_PTR(table); NRMS_REQUIRE_PTR(params);
    NRMS_REQUIRE_PTR(d_out); NRMS_REQUIRE_PTR(saved); NRMS_REQUIRE_PTR(scratch);
    NRMS_REQUIRE_PTR(d_params); NRMS_REQUIRE_PTR(d_rows);
    return encoder_bwd(*d, ids, table, params, d_out, saved, saved_bytes, scratch, scratch_bytes,
                       d_params, d_rows, true, 3, (cudaStream_t)stream);
}
int nrms_news_encoder_bwd_phase(const nrms_encoder_dims* d, const int64_t* ids, const float* table,
                                const float* params, const float* d_out, const void* saved,
                                int64_t saved_bytes, void* scratch, int64_t scratch_bytes,
                                float* d_params, float* d_rows, int32_t phase, nrms_stream_t stream) {
    int rc = check_dims(d, true);
    if (rc) return rc;
    if (phase != NRMS_BWD_DATA && phase != NRMS_BWD_PARAMS)
        return fail(NRMS_ERR_BAD_SHAPE, "phase=%d (expected NRMS_BWD_DATA or NRMS_BWD_PARAMS)", phase);
    NRMS_REQUIRE_PTR(ids); NRMS_REQUIRE_PTR(table); NRMS_REQUIRE_PTR(params);
    NRMS_REQUIRE_PTR(d_out); NRMS_REQUIRE_PTR(saved); NRMS_REQUIRE_PTR(scratch);
    NRMS_REQUIRE_PTR(d_params); NRMS_REQUIRE_PTR(d_rows);
    return encoder_bwd(*d, ids, table, params, d_out, saved, saved_bytes, scratch, scratch_bytes,
                       d_params, d_rows, true, phase, (cudaStream_t)stream);
}
int nrms_user_encoder_bwd_phase(const nrms_encoder_dims* d, const float* x, const float* params,
                                const float* d_out, const void* saved, int64_t saved_bytes, void* scratch,
                                int64_t scratch_bytes, float* d_params, float* d_x, int32_t phase,
                                nrms_stream_t stream) {
    int rc = check_dims(d, false);
    if (rc) return rc;
    if (phase != NRMS_BWD_DATA && phase != NRMS_BWD_PARAMS)
        return fail(NRMS_ERR_BAD_SHAPE, "phase=%d (expected NRMS_BWD_DATA or NRMS_BWD_PARAMS)", phase);
    NRMS_REQUIRE_PTR(x); NRMS_REQUIRE_PTR(params); NRMS_REQUIRE_PTR(d_out); NRMS_REQUIRE_PTR(saved);
    NRMS_REQUIRE_PTR(scratch); NRMS_REQUIRE_PTR(d_params); NRMS_REQUIRE_PTR(d_x);
    nrms_encoder_dims dd = *d;
    dd.dropout_p = 0.f;
    return encoder_bwd(dd, nullptr, x, params, d_out, saved, saved_bytes, scratch, scratch_bytes,
                       d_params, d_x, false, phase, (cudaStream_t)stream);
}
int nrms_user_encoder_fwd(const nrms_encoder_dims* d, const float* x, const float* params,
                          float* out, void* saved, int64_t saved_bytes, nrms_stream_t stream) {
    int rc = check_dims(d, false);
    if (rc) return rc;
    NRMS_REQUIRE_PTR(x); NRMS_REQUIRE_PTR(params); NRMS_REQUIRE_PTR(out); NRMS_REQUIRE_PTR(saved);
    nrms_encoder_dims dd = *d;
    dd.dropout_p = 0.f;  // UserEncoder has no dropout (nrms_v0.py:188-199)
    return encoder_fwd(dd, nullptr, x, params, out, saved, saved_bytes, false,
                       (cudaStream_t)stream);
}
int nrms_user_encoder_fwd_gather(const nrms_encoder_dims* d, const int64_t* ids, const float* table,
                                 const float* params, float* out, void* saved, int64_t saved_bytes,
                                 nrms_stream_t stream) {
    int rc = check_dims(d, true);
    if (rc) return rc;
    NRMS_REQUIRE_PTR(ids); NRMS_REQUIRE_PTR(table); NRMS_REQUIRE_PTR(params); NRMS_REQUIRE_PTR(out);
    NRMS_REQUIRE_PTR(saved);
    nrms_encoder_dims dd = *d;
    dd.dropout_p = 0.f;
    return encoder_fwd(dd, ids, table, params, out, saved, saved_bytes, false, (cudaStream_t)stream);
}
int nrms_user_encoder_bwd(const nrms_encoder_dims* d, const float* x, const float* params,
                          const float* d_out, const void* saved, int64_t saved_bytes,
                          void* scratch, int64_t scratch_bytes, float* d_params, float* d_x,
                          nrms_stream_t stream) {
    int rc = check_dims(d, false);
    if (rc) return rc;
    NRMS_REQUIRE_PTR(x); NRMS_REQUIRE_PTR(params); NRMS_REQUIRE_PTR(d_out);
    NRMS_REQUIRE_PTR(saved); NRMS_REQUIRE_PTR(scratch); NRMS_REQUIRE_PTR(d_params);
    NRMS_REQUIRE_PTR(d_x);
    nrms_encoder_dims dd = *d;
    dd.dropout_p = 0.f;
    return encoder_bwd(dd, nullptr, x, params, d_out, saved, saved_bytes, scratch, scratch_bytes,
                       d_params, d_x, false, 3, (cudaStream_t)stream);
}

static int score_check(int32_t B, int32_t C, int32_t D) {
    if (B < 1 || C < 1 || D < 1) return fail(NRMS_ERR_BAD_SHAPE, "B=%d C=%d D=%d", B, C, D);
    if (C > 8192) return fail(NRMS_ERR_BAD_SHAPE, "C=%d > 8192 unsupported", C);
    return NRMS_OK;
}

int nrms_score_fwd(int32_t B, int32_t C, int32_t D, const float* cand, const float* user,
                   const uint8_t* mask, float* logits, nrms_stream_t stream) {
    int rc = score_check(B, C, D);
    if (rc) return rc;
    NRMS_REQUIRE_PTR(cand); NRMS_REQUIRE_PTR(user); NRMS_REQUIRE_PTR(logits);
    ScoreArgs a{};
    a.cand = cand; a.user = user; a.mask = mask; a.logits = logits; a.B = B; a.C = C; a.D = D;
    NRMS_LAUNCH("score_0", (cudaStream_t)stream, score_kernel<0><<<B, 256, C * sizeof(float), (cudaStream_t)stream>>>(a));
    NRMS_CHECK_CUDA(cudaGetLastError());
    return NRMS_OK;
}
int nrms_score_cached(int32_t B, int32_t S, int32_t D, const float* vecs, int64_t n_vecs,
                      const int64_t* cand_ids, const float* user, const uint8_t* mask, float* logits,
                      nrms_stream_t stream) {
    int rc = score_check(B, S, D);
    if (rc) return rc;
    if (D % 4 || n_vecs < 1) return fail(NRMS_ERR_BAD_SHAPE, "D=%d must be a multiple of 4, n_vecs=%lld >= 1", D, (long long)n_vecs);
    NRMS_REQUIRE_PTR(vecs); NRMS_REQUIRE_PTR(user); NRMS_REQUIRE_PTR(logits);
    if (!cand_ids) return fail(NRMS_ERR_NULL, "cand_ids is NULL");
    ScoreCachedArgs a{vecs, n_vecs, cand_ids, user, mask, logits, B, S, D};
    NRMS_LAUNCH("score_cached", (cudaStream_t)stream,
                score_cached_kernel<<<B, 256, D * sizeof(float), (cudaStream_t)stream>>>(a));
    NRMS_CHECK_CUDA(cudaGetLastError());
    return NRMS_OK;
}
int nrms_score_bwd(int32_t B, int32_t C, int32_t D, const float* cand, const float* user,
                   const uint8_t* mask, const float* d_logits, float* d_cand, float* d_user,
                   nrms_stream_t stream) {
    int rc = score_check(B, C, D);
    if (rc) return rc;
    NRMS_REQUIRE_PTR(cand); NRMS_REQUIRE_PTR(user); NRMS_REQUIRE_PTR(d_logits);
    NRMS_REQUIRE_PTR(d_cand); NRMS_REQUIRE_PTR(d_user);
    ScoreArgs a{};
    a.cand = cand; a.user = user; a.mask = mask; a.d_logits = d_logits; a.d_cand = d_cand;
    a.d_user = d_user; a.B = B; a.C = C; a.D = D;
    NRMS_LAUNCH("score_2", (cudaStream_t)stream, score_kernel<2><<<B, 256, C * sizeof(float), (cudaStream_t)stream>>>(a));
    NRMS_CHECK_CUDA(cudaGetLastError());
    return NRMS_OK;
}
int nrms_score_ce_fwd_bwd(int32_t B, int32_t C, int32_t D, int32_t B_global, const float* cand,
                          const float* user, const uint8_t* mask, float* logits,
                          float* loss_per_row, float* d_cand, float* d_user,
                          float* loss_mean, uint32_t* ticket, nrms_stream_t stream) {
    int rc = score_check(B, C, D);
    if (rc) return rc;
    if (B_global < 1) return fail(NRMS_ERR_BAD_SHAPE, "B_global=%d", B_global);
    NRMS_REQUIRE_PTR(cand); NRMS_REQUIRE_PTR(user); NRMS_REQUIRE_PTR(logits);
    NRMS_REQUIRE_PTR(d_cand); NRMS_REQUIRE_PTR(d_user);
    if (!loss_per_row) return fail(NRMS_ERR_NULL, "loss_per_row is NULL");
    if (loss_mean && !ticket) return fail(NRMS_ERR_NULL, "ticket is NULL (required with loss_mean)");
    ScoreArgs a{};
    a.loss_mean = loss_mean; a.ticket = ticket;
    a.cand = cand; a.user = user; a.mask = mask; a.logits = logits; a.loss_rows = loss_per_row;
    a.d_cand = d_cand; a.d_user = d_user; a.B = B; a.C = C; a.D = D;
    a.inv_batch = 1.f / (float)B_global;
    NRMS_LAUNCH("score_1", (cudaStream_t)stream, score_kernel<1><<<B, 256, C * sizeof(float), (cudaStream_t)stream>>>(a));
    NRMS_CHECK_CUDA(cudaGetLastError());
    return NRMS_OK;
}

int64_t nrms_embedding_plan_bytes(int64_t n_rows, int32_t vocab) {
    if (n_rows < 0 || vocab < 1) return -1;
    return plan_bytes(n_rows, vocab);
}
int nrms_embedding_plan(const int64_t* ids, int64_t n_rows, int32_t vocab, void* plan,
                        int64_t plan_bytes_, nrms_stream_t stream) {
    if (n_rows < 1 || vocab < 1 || n_rows > 0x7fffffffll)
        return fail(NRMS_ERR_BAD_SHAPE, "n_rows=%lld vocab=%d", (long long)n_rows, vocab);
    NRMS_REQUIRE_PTR(ids); NRMS_REQUIRE_PTR(plan);
    if (plan_bytes_ < plan_bytes(n_rows, vocab))
        return fail(NRMS_ERR_WORKSPACE, "plan blob %lld < %lld bytes", (long long)plan_bytes_,
                    (long long)plan_bytes(n_rows, vocab));
    cudaStream_t s = (cudaStream_t)stream;
    PlanView v = plan_view(plan, n_rows, vocab);
    // 1. counts -> offsets, n_valid
    NRMS_CHECK_CUDA(cudaMemsetAsync(v.counts, 0, sizeof(int32_t) * vocab, s));
    NRMS_LAUNCH("plan_hist", s, plan_hist_kernel<<<grid_for(n_rows, 256), 256, 0, s>>>(ids, n_rows, vocab, v.counts));
    const int scan_blocks = ceil_div(vocab, kScanBlock);
    NRMS_LAUNCH("plan_scan", s, plan_block_totals_kernel<<<scan_blocks, kScanBlock, 0, s>>>(v.counts, v.block_tot, vocab));
    NRMS_LAUNCH("plan_scan", s, plan_scan_kernel<<<scan_blocks, kScanBlock, 0, s>>>(v.counts, v.block_tot, v.offsets, v.n_valid, vocab));
    // 2. stable LSD radix sort of (id, row): keys run up to `vocab` (the key of a padding id)
    int bits = 1;
    while ((1ll << bits) <= (long long)vocab) ++bits;
    const int rb = bits <= 18 ? 9 : 8, passes = ceil_div(bits, rb), bins = 1 << rb;
    const int nblk = (int)ceil_div64(n_rows, kRsBlock);
    int32_t *ka = v.sorted_id, *va = v.perm, *kb = v.tmp_key, *vb = v.tmp_val;
    if (passes % 2) { std::swap(ka, kb); std::swap(va, vb); }   // an odd number of passes must end in (sorted_id, perm)
    NRMS_LAUNCH("plan_sort", s, rsort_init_kernel<<<grid_for(n_rows, 256), 256, 0, s>>>(ids, n_rows, vocab, ka, va));
    for (int p = 0; p < passes; ++p) {
        NRMS_LAUNCH("plan_sort", s, rsort_hist_kernel<<<nblk, 256, 0, s>>>(ka, n_rows, p * rb, bins, v.rhist, nblk));
        const int nh = bins * nblk, hblocks = ceil_div(nh, kScanBlock);
        NRMS_LAUNCH("plan_sort", s, plan_block_totals_kernel<<<hblocks, kScanBlock, 0, s>>>(v.rhist, v.rblock_tot, nh));
        NRMS_LAUNCH("plan_sort", s, plan_scan_kernel<<<hblocks, kScanBlock, 0, s>>>(v.rhist, v.rblock_tot, v.rscan, nullptr, nh));
        NRMS_LAUNCH("plan_sort", s, rsort_scatter_kernel<<<nblk, 256, 0, s>>>(ka, va, n_rows, p * rb, bins, v.rscan, nblk, kb, vb));
        std::swap(ka, kb);
        std::swap(va, vb);
    }
    NRMS_CHECK_CUDA(cudaGetLastError());
    return NRMS_OK;
}
int nrms_embedding_grad_dense(const void* plan, int64_t plan_bytes_, const float* d_rows,
                              int64_t n_rows, int32_t vocab, int32_t D, float* d_table,
                              nrms_stream_t stream) {
    if (n_rows < 1 || vocab < 1 || D < 4 || D % 4 || D > 384)
        return fail(NRMS_ERR_BAD_SHAPE, "n_rows=%lld vocab=%d D=%d", (long long)n_rows, vocab, D);
    NRMS_REQUIRE_PTR(plan); NRMS_REQUIRE_PTR(d_rows); NRMS_REQUIRE_PTR(d_table);
    if (plan_bytes_ < plan_bytes(n_rows, vocab))
        return fail(NRMS_ERR_WORKSPACE, "plan blob too small");
    cudaStream_t s = (cudaStream_t)stream;
    PlanView v = plan_view(const_cast<void*>(plan), n_rows, vocab);
    const long long n4 = (long long)vocab * D / 4;
    NRMS_LAUNCH("zero", s, zero_kernel<<<grid_for(n4, 256), 256, 0, s>>>(reinterpret_cast<float4*>(d_table), n4));
    NRMS_LAUNCH("embgrad_reduce", s, embgrad_reduce_kernel<<<grid_for(n_rows, 256, 16), 256, 0, s>>>(
        v.perm, v.sorted_id, v.offsets, v.n_valid, d_rows, D, d_table, v.part, v.n_chunks));
    NRMS_LAUNCH("embgrad_reduce", s, embgrad_fixup_kernel<<<grid_for((long long)vocab * 32, 256, 8), 256, 0, s>>>(
        v.offsets, vocab, D, v.part, v.n_chunks, d_table));
    NRMS_CHECK_CUDA(cudaGetLastError());
    return NRMS_OK;
}
int nrms_embedding_plan_unique(const void* plan, int64_t plan_bytes_, int32_t vocab,
                               int32_t* d_unique, nrms_stream_t stream) {
    NRMS_REQUIRE_PTR(plan);
    if (!d_unique) return fail(NRMS_ERR_NULL, "d_unique is NULL");
    (void)plan_bytes_;
    cudaStream_t s = (cudaStream_t)stream;
    NRMS_CHECK_CUDA(cudaMemsetAsync(d_unique, 0, sizeof(int32_t), s));
    NRMS_LAUNCH("plan_unique", s, plan_unique_kernel<<<grid_for(vocab, 256), 256, 0, s>>>(
        reinterpret_cast<const int32_t*>(plan), vocab, d_unique));
    NRMS_CHECK_CUDA(cudaGetLastError());
    return NRMS_OK;
}

int nrms_adam_step(float* p, const float* g, float* m, float* v, int64_t n, int32_t step,
                   float lr, float beta1, float beta2, float eps, float grad_scale,
                   nrms_stream_t stream) {
    if (n < 1 || step < 1) return fail(NRMS_ERR_BAD_SHAPE, "n=%lld step=%d", (long long)n, step);
    NRMS_REQUIRE_PTR(p); NRMS_REQUIRE_PTR(g); NRMS_REQUIRE_PTR(m); NRMS_REQUIRE_PTR(v);
    AdamArgs a{};
    a.p = p; a.g = g; a.m = m; a.v = v; a.n = n;
    a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.grad_scale = grad_scale;
    // bias corrections exactly as torch/optim/adam.py (_single_tensor_adam, python floats)
    const double bc1 = 1.0 - pow((double)beta1, (double)step);
    const double bc2 = 1.0 - pow((double)beta2, (double)step);
    a.step_size = (float)((double)lr / bc1);
    a.bc2_sqrt = (float)sqrt(bc2);
    NRMS_LAUNCH("adam", (cudaStream_t)stream, adam_kernel<<<grid_for(n / 4 + 1, 256), 256, 0, (cudaStream_t)stream>>>(a));
    NRMS_CHECK_CUDA(cudaGetLastError());
    return NRMS_OK;
}

static int metrics_impl(const float* scores, long long row_stride, const uint8_t* labels,
                        const int64_t* offsets, int64_t n_impr, int32_t max_len, double* out,
                        cudaStream_t s) {
    if (n_impr < 1 || max_len < 1 || max_len > 8192)
        return fail(NRMS_ERR_BAD_SHAPE, "n_impr=%lld max_len=%d", (long long)n_impr, max_len);
    if (!scores || !labels || !offsets || !out) return fail(NRMS_ERR_NULL, "NULL argument");
    const size_t smem = (size_t)kMetricWarps * max_len * (sizeof(float) + 1);
    NRMS_CHECK_CUDA(cudaFuncSetAttribute(rank_metrics_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = grid_for(n_impr * 32, kMetricWarps * 32, 8);
    NRMS_LAUNCH("rank_metrics", s, rank_metrics_kernel<<<grid, kMetricWarps * 32, smem, s>>>(scores, row_stride, labels, offsets,
                                                             n_impr, max_len, out));
    NRMS_CHECK_CUDA(cudaGetLastError());
    return NRMS_OK;
}
int nrms_rank_metrics(const float* scores, const uint8_t* labels, const int64_t* offsets,
                      int64_t n_impr, int32_t max_len, double* out, nrms_stream_t stream) {
    return metrics_impl(scores, -1, labels, offsets, n_impr, max_len, out, (cudaStream_t)stream);
}
int nrms_rank_metrics_padded(const float* scores, int64_t row_stride, const uint8_t* labels,
                             const int64_t* offsets, int64_t n_impr, int32_t max_len,
                             double* out, nrms_stream_t stream) {
    if (row_stride < max_len)
        return fail(NRMS_ERR_BAD_SHAPE, "row_stride=%lld < max_len=%d", (long long)row_stride,
                    max_len);
    return metrics_impl(scores, row_stride, labels, offsets, n_impr, max_len, out,
                        (cudaStream_t)stream);
}

int nrms_rank_metrics_rows(const float* scores, int64_t row_stride, const uint8_t* labels,
                           int64_t label_stride, const int64_t* lens, int64_t n_impr, int32_t max_len,
                           double* out, nrms_stream_t stream) {
    if (n_impr < 1 || max_len < 1 || max_len > 8192 || row_stride < max_len || label_stride < max_len)
        return fail(NRMS_ERR_BAD_SHAPE, "n_impr=%lld max_len=%d row_stride=%lld label_stride=%lld", (long long)n_impr,
                    max_len, (long long)row_stride, (long long)label_stride);
    if (!scores || !labels || !lens || !out) return fail(NRMS_ERR_NULL, "NULL argument");
    cudaStream_t s = (cudaStream_t)stream;
    const size_t smem = (size_t)kMetricWarps * max_len * (sizeof(float) + 1);
    NRMS_CHECK_CUDA(cudaFuncSetAttribute(rank_metrics_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = grid_for(n_impr * 32, kMetricWarps * 32, 8);
    NRMS_LAUNCH("rank_metrics", s, rank_metrics_kernel<<<grid, kMetricWarps * 32, smem, s>>>(
        scores, row_stride, labels, nullptr, n_impr, max_len, out, label_stride, lens));
    NRMS_CHECK_CUDA(cudaGetLastError());
    return NRMS_OK;
}

int nrms_assemble_batch(const int64_t* index, int32_t B, const int64_t* browsed_ids,
                        const int64_t* browsed_lens, const int64_t* candidate_ids,
                        const int64_t* candidate_lens, const int64_t* titles, int64_t n_news,
                        int32_t H, int32_t S, int32_t T, int64_t* o_browsed_ids,
                        int64_t* o_browsed_lens, int64_t* o_browsed_titles, uint8_t* o_browsed_mask,
                        int64_t* o_candidate_ids, int64_t* o_candidate_titles,
                        uint8_t* o_candidate_mask, nrms_stream_t stream) {
    if (B < 1 || H < 1 || S < 1 || T < 1 || n_news < 1)
        return fail(NRMS_ERR_BAD_SHAPE, "B=%d H=%d S=%d T=%d n_news=%lld", B, H, S, T, (long long)n_news);
    // 8-byte elements, scalar accesses: no 16-byte alignment requirement (index is usually a slice)
    if (!index || !browsed_ids || !browsed_lens || !candidate_ids || !candidate_lens || !titles || !o_browsed_ids ||
        !o_browsed_lens || !o_browsed_titles || !o_browsed_mask || !o_candidate_ids || !o_candidate_titles ||
        !o_candidate_mask)
        return fail(NRMS_ERR_NULL, "NULL argument");
    AssembleArgs a{index, browsed_ids, browsed_lens, candidate_ids, candidate_lens, titles, n_news, B, H, S, T,
                   o_browsed_ids, o_browsed_lens, o_browsed_titles, o_browsed_mask, o_candidate_ids,
                   o_candidate_titles, o_candidate_mask};
    cudaStream_t s = (cudaStream_t)stream;
    NRMS_LAUNCH("assemble_batch", s, assemble_batch_kernel<<<grid_for((long long)B * (H + S) * 32, 256, 16), 256, 0, s>>>(a));
    NRMS_CHECK_CUDA(cudaGetLastError());
    return NRMS_OK;
}

int nrms_rank_positions(const float* scores, int64_t row_stride, const int64_t* lens,
                        int64_t n_impr, int32_t* ranks, nrms_stream_t stream) {
    if (n_impr < 1 || row_stride < 1 || row_stride > 4096)
        return fail(NRMS_ERR_BAD_SHAPE, "n_impr=%lld row_stride=%lld", (long long)n_impr, (long long)row_stride);
    NRMS_REQUIRE_PTR(scores); NRMS_REQUIRE_PTR(lens); NRMS_REQUIRE_PTR(ranks);
    cudaStream_t s = (cudaStream_t)stream;
    const size_t smem = (size_t)kMetricWarps * row_stride * sizeof(float);
    int rc = set_smem(rank_positions_kernel, smem);
    if (rc) return rc;
    const unsigned grid = (unsigned)std::min<long long>(ceil_div64(n_impr, kMetricWarps), 8ll * kNumSMs);
    NRMS_LAUNCH("rank_positions", s, rank_positions_kernel<<<grid, kMetricWarps * 32, smem, s>>>(
        scores, row_stride, lens, n_impr, (int)row_stride, ranks));
    NRMS_CHECK_CUDA(cudaGetLastError());
    return NRMS_OK;
}

int nrms_gather_rows_f32(const float* src, int64_t n_src, int32_t D, const int64_t* idx,
                         int64_t n_idx, int64_t base, float* out, nrms_stream_t stream) {
    if (n_src < 1 || D < 1 || n_idx < 1) return fail(NRMS_ERR_BAD_SHAPE, "bad gather shape");
    if (!src || !idx || !out) return fail(NRMS_ERR_NULL, "NULL argument");
    NRMS_LAUNCH("gather_rows_float", (cudaStream_t)stream, gather_rows_kernel<float><<<grid_for(n_idx * 32, 256, 16), 256, 0, (cudaStream_t)stream>>>(
        src, n_src, D, idx, n_idx, base, out));
    NRMS_CHECK_CUDA(cudaGetLastError());
    return NRMS_OK;
}
int nrms_gather_rows_i64(const int64_t* src, int64_t n_src, int32_t D, const int64_t* idx,
                         int64_t n_idx, int64_t base, int64_t* out, nrms_stream_t stream) {
    if (n_src < 1 || D < 1 || n_idx < 1) return fail(NRMS_ERR_BAD_SHAPE, "bad gather shape");
    if (!src || !idx || !out) return fail(NRMS_ERR_NULL, "NULL argument");
    NRMS_LAUNCH("gather_rows_int64_t", (cudaStream_t)stream, gather_rows_kernel<int64_t><<<grid_for(n_idx * 32, 256, 16), 256, 0, (cudaStream_t)stream>>>(
        src, n_src, D, idx, n_idx, base, out));
    NRMS_CHECK_CUDA(cudaGetLastError());
    return NRMS_OK;
}

int nrms_dropout_mask(uint64_t seed, uint32_t stream_id, float p, int64_t n_rows, int32_t n_cols,
                      float* out, nrms_stream_t stream) {
    if (n_rows < 1 || n_cols < 1 || p < 0.f || p >= 1.f)
        return fail(NRMS_ERR_BAD_SHAPE, "n_rows=%lld n_cols=%d p=%f", (long long)n_rows, n_cols, (double)p);
    if (!out) return fail(NRMS_ERR_NULL, "out is NULL");
    NRMS_LAUNCH("dropout_mask", (cudaStream_t)stream,
                dropout_mask_kernel<<<grid_for(n_rows * ceil_div(n_cols, 8), 256), 256, 0, (cudaStream_t)stream>>>(
                    make_dropout(p, seed), stream_id, n_rows, n_cols, out));
    NRMS_CHECK_CUDA(cudaGetLastError());
    return NRMS_OK;
}

// ---- self-test of the tcgen05 GEMM layer (tests/test_gpu_gemm.py) ------------------------------
namespace {
struct SelfTestLayout {
    ig::Img a, b;
    float* bias;
    float* part;
    int splits, a_rows, a_chunks, b_rows, b_chunks;
    int64_t bytes;
};
SelfTestLayout selftest_layout(void* blob, int variant, int M, int N, int K) {
    SelfTestLayout L{};
    char* p = reinterpret_cast<char*>(blob);
    int64_t off = 0;
    auto take = [&](int64_t n) {
        char* r = p + off;
        off += align_up(n, 1024);
        return r;
    };
    if (variant == 0 || variant == 3) {
        const int nt = variant == 0 ? 240 : 256;
        L.a_rows = M; L.a_chunks = ceil_div(K, 64);
        L.b_rows = ceil_div(N, nt) * nt; L.b_chunks = L.a_chunks;
    } else if (variant == 1) {
        L.a_rows = M; L.a_chunks = ceil_div(K, 64);
        L.b_rows = K; L.b_chunks = 5;
    } else {
        L.a_rows = K; L.a_chunks = 2 * ceil_div(M, 128);
        L.b_rows = K; L.b_chunks = 5;
    }
    L.a = ig::img_view(take(ig::img_bytes(L.a_rows, L.a_chunks)), L.a_rows, L.a_chunks);
    L.b = ig::img_view(take(ig::img_bytes(L.b_rows, L.b_chunks)), L.b_rows, L.b_chunks);
    L.bias = reinterpret_cast<float*>(take((int64_t)N * 4));
    L.splits = variant == 2 ? wgrad_splits_tc(ceil_div(M, 128), L.a.rows_pad / 64) : 1;
    L.part = reinterpret_cast<float*>(take(variant == 2 ? (int64_t)L.splits * M * N * 4 : 0));
    L.bytes = off;
    return L;
}
}  // namespace

int64_t nrms_gemm_selftest_bytes(int32_t variant, int32_t M, int32_t N, int32_t K) {
    if (variant < 0 || variant > 3 || M < 1 || N < 1 || K < 1) return -1;
    return selftest_layout(nullptr, variant, M, N, K).bytes;
}
int nrms_gemm_selftest(int32_t variant, const float* A, const float* B, float* C, int32_t M,
                       int32_t N, int32_t K, void* work, int64_t work_bytes, nrms_stream_t stream) {
    if (variant < 0 || variant > 3 || M < 1 || N < 1 || K < 1 || N % 4)
        return fail(NRMS_ERR_BAD_SHAPE, "variant=%d M=%d N=%d K=%d", variant, M, N, K);
    if (variant != 0 && variant != 3 && N > 320) return fail(NRMS_ERR_BAD_SHAPE, "N=%d > 320 for variant %d", N, variant);
    NRMS_REQUIRE_PTR(A); NRMS_REQUIRE_PTR(B); NRMS_REQUIRE_PTR(C); NRMS_REQUIRE_PTR(work);
    const SelfTestLayout L = selftest_layout(work, variant, M, N, K);
    if (work_bytes < L.bytes) return fail(NRMS_ERR_WORKSPACE, "work blob %lld < %lld", (long long)work_bytes, (long long)L.bytes);
    cudaStream_t s = (cudaStream_t)stream;
    if (variant == 0) {
        NRMS_CHECK_CUDA(ig::img_pack(A, M, K, K, L.a, s));
        NRMS_CHECK_CUDA(ig::img_pack(B, N, K, K, L.b, s));
        NRMS_CHECK_CUDA(cudaMemsetAsync(L.bias, 0, (size_t)N * 4, s));
        ig::IgArgs g = ig_args(L.a, L.b, C, N, M, N);
        g.bias = L.bias;
        g.m_tiles = L.a.rows_pad / 128; g.n_tiles = ceil_div(N, 240);
        g.k_steps = ceil_div(K, 16); g.k_chunks = ceil_div(g.k_steps, 4);
        NRMS_CHECK_CUDA((ig::ig_launch<false, false, 240, ig::EPI_BIAS>(g, s, "selftest_nt")));
    } else if (variant == 3) {
        // variant 0's contraction on CTA pairs (cta_group::2, M = 256 per pair)
        NRMS_CHECK_CUDA(ig::img_pack(A, M, K, K, L.a, s));
        NRMS_CHECK_CUDA(ig::img_pack(B, N, K, K, L.b, s));
        NRMS_CHECK_CUDA(cudaMemsetAsync(L.bias, 0, (size_t)N * 4, s));
        ig::IgArgs g = ig_args(L.a, L.b, C, N, M, N);
        g.bias = L.bias;
        g.m_tiles = L.a.rows_pad / 128; g.n_tiles = ceil_div(N, 256);
        g.k_steps = ceil_div(K, 16); g.k_chunks = ceil_div(g.k_steps, 4);
        NRMS_CHECK_CUDA((ig::ig_launch<false, false, 256, ig::EPI_BIAS, true>(g, s, "selftest_nt_pair")));
    } else if (variant == 1) {
        NRMS_CHECK_CUDA(ig::img_pack(A, M, K, K, L.a, s));
        NRMS_CHECK_CUDA(ig::img_pack(B, K, N, N, L.b, s));
        ig::IgArgs g = ig_args(L.a, L.b, C, N, M, N);
        g.m_tiles = L.a.rows_pad / 128; g.n_tiles = 1;
        g.k_steps = ceil_div(K, 16); g.k_chunks = ceil_div(g.k_steps, 4);
        NRMS_CHECK_CUDA((ig::ig_launch<false, true, 320, ig::EPI_MASK>(g, s, "selftest_nn")));
    } else {
        NRMS_CHECK_CUDA(ig::img_pack(A, K, M, M, L.a, s));
        NRMS_CHECK_CUDA(ig::img_pack(B, K, N, N, L.b, s));
        ig::IgArgs g = ig_args(L.a, L.b, L.part, N, M, N);
        g.m_tiles = ceil_div(M, 128); g.n_tiles = 1;
        g.k_chunks = L.a.rows_pad / 64; g.k_steps = 4 * g.k_chunks;
        g.splits = L.splits;
        g.c_split_stride = (long long)M * N;
        NRMS_CHECK_CUDA((ig::ig_launch<true, true, 320, ig::EPI_PARTIAL>(g, s, "selftest_tn")));
        int rc = reduce_rows(L.part, C, g.splits, (long long)M * N, (long long)M * N, 1.f, nullptr, s);
        if (rc) return rc;
    }
    return NRMS_OK;
}

int nrms_validate_ids(const int64_t* ids, int64_t n, int64_t vocab, int32_t* d_flag,
                      nrms_stream_t stream) {
    if (n < 1) return fail(NRMS_ERR_BAD_SHAPE, "n=%lld", (long long)n);
    if (!ids || !d_flag) return fail(NRMS_ERR_NULL, "NULL argument");
    NRMS_LAUNCH("validate_ids", (cudaStream_t)stream, validate_ids_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(ids, n, vocab, d_flag));
    NRMS_CHECK_CUDA(cudaGetLastError());
    return NRMS_OK;
}

#include "abi_masked.inc"

}  // extern "C"
