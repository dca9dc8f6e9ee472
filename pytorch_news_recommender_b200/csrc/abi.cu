// abi.cu — the extern "C" surface of libnrms_b200.so (see include/nrms_b200.h) and the
// orchestration of the kernels behind each entry point.  No allocation, no sync: every
// call only enqueues kernels on the caller's stream.
#include "../../include/nrms_b200.h"

#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "attention.cuh"
#include "common.cuh"
#include "gemm_simt.cuh"
#include "gemm_tc.cuh"
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
struct Saved {
    float* qkv;  // [M, 3D]
    float* lse;  // [M, h]
    float* ctx;  // [M, D]   (post-dropout)
    float* t;    // [M, Q]
    float* w;    // [n_seq, L]
    // tcgen05 path: weights pre-split to bf16 hi/lo in the swizzled smem image (gemm_tc.cuh)
    uint8_t* pk_qkv;    // B = W_qkv   [3D, D]   (forward projections)
    uint8_t* pk_a;      // B = W_a     [Q, D]    (forward additive projection)
    uint8_t* pk_qkv_t;  // B = W_qkv^T [D, 3D]   (data gradient)
    uint8_t* pk_a_t;    // B = W_a^T   [D, Q]    (data gradient)
    int64_t bytes;
};
Saved saved_layout(void* blob, const nrms_encoder_dims& d) {
    const int64_t M = (int64_t)d.n_seq * d.seq_len;
    char* p = reinterpret_cast<char*>(blob);
    int64_t off = 0;
    auto take = [&](int64_t nfloat) {
        float* r = reinterpret_cast<float*>(p + off);
        off += align_up(nfloat * (int64_t)sizeof(float), 256);
        return r;
    };
    Saved s;
    s.qkv = take(M * 3 * d.d_model);
    s.lse = take(M * d.n_heads);
    s.ctx = take(M * d.d_model);
    s.t = take(M * d.d_query);
    s.w = take((int64_t)d.n_seq * d.seq_len);
    auto take_bytes = [&](int64_t n) {
        uint8_t* r = reinterpret_cast<uint8_t*>(p + off);
        off += align_up(n, 1024);
        return r;
    };
    const int D = d.d_model, Q = d.d_query;
    s.pk_qkv = take_bytes(tc::packed_b_bytes(3 * D, D, tc::pick_n_tile(3 * D)));
    s.pk_a = take_bytes(tc::packed_b_bytes(Q, D, tc::pick_n_tile(Q)));
    s.pk_qkv_t = take_bytes(tc::packed_b_bytes(D, 3 * D, tc::pick_n_tile(D)));
    s.pk_a_t = take_bytes(tc::packed_b_bytes(D, Q, tc::pick_n_tile(D)));
    s.bytes = off;
    return s;
}

constexpr int kMaxSplits = 32;
constexpr int kReduceSlices = 128;

int wgrad_splits(int out_rows, int out_cols, long long k_rows) {
    const int tiles = ceil_div(out_rows, GBM) * ceil_div(out_cols, GBN);
    int s = ceil_div(4 * kNumSMs, tiles);
    if (s > kMaxSplits) s = kMaxSplits;
    const long long max_by_k = ceil_div64(k_rows, 4 * GBK);
    if (s > max_by_k) s = (int)max_by_k;
    return s < 1 ? 1 : s;
}

struct Scratch {
    float* d_ctx;      // [M, D]
    float* d_pre;      // [M, Q]
    float* d_qkv;      // [M, 3D]
    float* part_q;     // [n_seq, 2Q]
    float* part_b;     // [n_seq, 3D]
    float* wpart;      // [kMaxSplits, 3D*D]
    float* red_tmp;    // [kReduceSlices, max(3D, 2Q)]
    int64_t bytes;
};
Scratch scratch_layout(void* blob, const nrms_encoder_dims& d) {
    const int64_t M = (int64_t)d.n_seq * d.seq_len;
    const int64_t D = d.d_model, Q = d.d_query;
    char* p = reinterpret_cast<char*>(blob);
    int64_t off = 0;
    auto take = [&](int64_t nfloat) {
        float* r = reinterpret_cast<float*>(p + off);
        off += align_up(nfloat * (int64_t)sizeof(float), 256);
        return r;
    };
    Scratch s;
    s.d_ctx = take(M * D);
    s.d_pre = take(M * Q);
    s.d_qkv = take(M * 3 * D);
    s.part_q = take((int64_t)d.n_seq * 2 * Q);
    s.part_b = take((int64_t)d.n_seq * 3 * D);
    s.wpart = take((int64_t)kMaxSplits * 3 * D * (D > Q ? D : Q));
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
    if (d->gemm_mode != 0 && d->gemm_mode != 1)
        return fail(NRMS_ERR_BAD_SHAPE, "gemm_mode=%d unknown", d->gemm_mode);
    return NRMS_OK;
}

// deterministic column sums: out[n] = scale * sum_r in[r, n]
int reduce_rows(const float* in, float* out, long long R, long long n, long long ld, float scale,
                float* tmp, cudaStream_t s) {
    const int threads = 256;
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

int pick_hpb(int L, int n_heads, bool bwd) {
    const int rc = ceil_div(L, 32);
    const int max_warps = bwd ? 8 : 16;
    const size_t budget = 100 * 1024;
    int hpb = n_heads;
    while (hpb > 1 && (hpb * rc > max_warps ||
                       (bwd ? attn_bwd_smem_bytes(L, hpb) : attn_fwd_smem_bytes(L, hpb)) > budget))
        --hpb;
    return hpb;
}

// y[M,N] = epi(x W^T + b): dispatches on gemm_mode
int linear_fwd(const nrms_encoder_dims& d, const float* x, const int64_t* gather_rows, int M,
               int N, int K, const float* W, const float* bias, float* y, int epilogue,
               bool drop_in, uint8_t* packed, cudaStream_t s) {
    GemmArgs g{};
    g.A = x; g.B = W; g.C = y; g.bias = bias; g.a_rows = gather_rows; g.b_rows = nullptr;
    g.M = M; g.N = N; g.K = K; g.lda = K; g.ldb = K; g.ldc = N;
    g.k_chunk = K; g.c_split_stride = 0; g.accumulate = 0; g.epilogue = epilogue;
    g.drop = make_dropout(d.dropout_p, d.seed);
    g.drop_on = (drop_in && g.drop.enabled()) ? 1 : 0;
    g.drop_sid = kDropEmbedding;
    const char* name = N == 3 * K ? "gemm_fwd_qkv" : "gemm_fwd_additive";
    if (d.gemm_mode == 1) {
        NRMS_CHECK_CUDA(tc::pack_b(W, N, K, K, 1, packed, s));
        NRMS_CHECK_CUDA(tc::launch(g, packed, 3, s, name));
        return NRMS_OK;
    }
    NRMS_CHECK_CUDA(launch_gemm_simt(g, true, true, 1, s, name));
    return NRMS_OK;
}

int encoder_fwd(const nrms_encoder_dims& d, const int64_t* ids, const float* x_or_table,
                const float* params, float* out, void* saved_blob, int64_t saved_bytes,
                bool news, cudaStream_t s) {
    const Saved sv = saved_layout(saved_blob, d);
    if (saved_bytes < sv.bytes)
        return fail(NRMS_ERR_WORKSPACE, "saved blob %lld < %lld bytes", (long long)saved_bytes,
                    (long long)sv.bytes);
    const int D = d.d_model, Q = d.d_query, L = d.seq_len, h = d.n_heads, dk = D / h;
    const int M = d.n_seq * L;
    const ParamView pv = param_view<ParamView>(params, D, Q);
    const Dropout drop = make_dropout(d.dropout_p, d.seed);

    // 1. Q|K|V projections (nrms_v0.py:53-58) with the embedding gather + dropout fused into
    //    the A-operand load (nrms_v0.py:166)
    int rc = linear_fwd(d, x_or_table, news ? ids : nullptr, M, 3 * D, D, pv.Wqkv, pv.bqkv,
                        sv.qkv, 0, news, sv.pk_qkv, s);
    if (rc) return rc;
    // 2. per-head attention (+ context dropout for the news encoder)
    {
        AttnArgs a{};
        a.qkv = sv.qkv; a.ctx = sv.ctx; a.lse = sv.lse;
        a.L = L; a.D = D; a.n_heads = h; a.dk = dk;
        a.hpb = pick_hpb(L, h, false);
        a.scale = 1.f / sqrtf((float)dk);
        a.drop = news ? drop : make_dropout(0.f, 0);
        const size_t smem = attn_fwd_smem_bytes(L, a.hpb);
        NRMS_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem));
        const int warps = a.hpb * ceil_div(L, 32);
        NRMS_LAUNCH("attn_fwd", s, attn_fwd_kernel<<<dim3(d.n_seq, ceil_div(h, a.hpb)), warps * 32, smem, s>>>(a));
        NRMS_CHECK_CUDA(cudaGetLastError());
    }
    // 3. additive-attention projection t = tanh(ctx W_a^T + b_a) (nrms_v0.py:108)
    rc = linear_fwd(d, sv.ctx, nullptr, M, Q, D, pv.Wa, pv.ba, sv.t, 1, false, sv.pk_a, s);
    if (rc) return rc;
    // 4. softmax over the sequence + weighted sum (nrms_v0.py:110-126)
    {
        PoolArgs p{};
        p.ctx = sv.ctx; p.t = sv.t; p.q = pv.qv; p.w = sv.w; p.out = out;
        p.L = L; p.D = D; p.Q = Q;
        NRMS_LAUNCH("pool_fwd", s, pool_fwd_kernel<<<d.n_seq, 256, L * sizeof(float), s>>>(p));
        NRMS_CHECK_CUDA(cudaGetLastError());
    }
    return NRMS_OK;
}

int encoder_bwd(const nrms_encoder_dims& d, const int64_t* ids, const float* x_or_table,
                const float* params, const float* d_out, const void* saved_blob,
                int64_t saved_bytes, void* scratch_blob, int64_t scratch_bytes, float* d_params,
                float* d_x, bool news, cudaStream_t s) {
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
    const ParamView pv = param_view<ParamView>(params, D, Q);
    const GradView gv = param_view<GradView>(d_params, D, Q);
    const Dropout drop = make_dropout(d.dropout_p, d.seed);
    int rc;

    // 1. pooling backward: d_ctx (pool path), d_pre, partials of d_b_a and d_query
    {
        PoolArgs p{};
        p.ctx = sv.ctx; p.t = sv.t; p.q = pv.qv; p.w = sv.w; p.d_out = d_out;
        p.d_ctx = sc.d_ctx; p.d_pre = sc.d_pre; p.d_part = sc.part_q;
        p.L = L; p.D = D; p.Q = Q;
        NRMS_LAUNCH("pool_bwd", s, pool_bwd_kernel<<<d.n_seq, 256, 2 * L * sizeof(float), s>>>(p));
        NRMS_CHECK_CUDA(cudaGetLastError());
    }
    // [d_b_a | d_query] are adjacent in the flat block, as in part_q
    rc = reduce_rows(sc.part_q, gv.ba, d.n_seq, 2 * Q, 2 * Q, 1.f, sc.red_tmp, s);
    if (rc) return rc;
    // 2. d_ctx += d_pre W_a
    {
        GemmArgs g{};
        g.A = sc.d_pre; g.B = pv.Wa; g.C = sc.d_ctx;
        g.M = M; g.N = D; g.K = Q; g.lda = Q; g.ldb = D; g.ldc = D;
        g.k_chunk = Q; g.accumulate = 1;
        if (d.gemm_mode == 1) {
            // B(n = d, k = q) = W_a[q, d]
            NRMS_CHECK_CUDA(tc::pack_b(pv.Wa, D, Q, 1, D, sv.pk_a_t, s));
            NRMS_CHECK_CUDA(tc::launch(g, sv.pk_a_t, 3, s, "gemm_dgrad_additive"));
        } else {
            NRMS_CHECK_CUDA(launch_gemm_simt(g, true, false, 1, s, "gemm_dgrad_additive"));
        }
    }
    // 3. dW_a = d_pre^T ctx   (reduction over the M token rows, split + deterministic reduce)
    {
        const int splits = wgrad_splits(Q, D, M);
        GemmArgs g{};
        g.A = sc.d_pre; g.B = sv.ctx; g.C = sc.wpart;
        g.M = Q; g.N = D; g.K = M; g.lda = Q; g.ldb = D; g.ldc = D;
        g.k_chunk = (int)align_up(ceil_div(M, splits), GBK);
        g.c_split_stride = (long long)Q * D;
        NRMS_CHECK_CUDA(launch_gemm_simt(g, false, false, ceil_div(M, g.k_chunk), s, "gemm_wgrad_additive"));
        rc = reduce_rows(sc.wpart, gv.Wa, ceil_div(M, g.k_chunk), (long long)Q * D,
                         (long long)Q * D, 1.f, nullptr, s);
        if (rc) return rc;
    }
    // 4. attention backward -> d_qkv and bias partials
    {
        AttnArgs a{};
        a.qkv = sv.qkv; a.ctx = sv.ctx; a.lse = sv.lse; a.d_ctx = sc.d_ctx; a.d_qkv = sc.d_qkv;
        a.d_bias_part = sc.part_b;
        a.L = L; a.D = D; a.n_heads = h; a.dk = dk;
        a.hpb = pick_hpb(L, h, true);
        a.scale = 1.f / sqrtf((float)dk);
        a.drop = news ? drop : make_dropout(0.f, 0);
        const size_t smem = attn_bwd_smem_bytes(L, a.hpb);
        NRMS_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_kernel,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem));
        const int warps = a.hpb * ceil_div(L, 32);
        NRMS_LAUNCH("attn_bwd", s, attn_bwd_kernel<<<dim3(d.n_seq, ceil_div(h, a.hpb)), warps * 32, smem, s>>>(a));
        NRMS_CHECK_CUDA(cudaGetLastError());
    }
    rc = reduce_rows(sc.part_b, gv.bqkv, d.n_seq, 3 * D, 3 * D, 1.f, sc.red_tmp, s);
    if (rc) return rc;
    // 5. dW_qkv = d_qkv^T x   (x = gathered + dropped embedding rows for the news encoder)
    {
        const int splits = wgrad_splits(3 * D, D, M);
        GemmArgs g{};
        g.A = sc.d_qkv; g.B = x_or_table; g.C = sc.wpart;
        g.b_rows = news ? ids : nullptr;
        g.M = 3 * D; g.N = D; g.K = M; g.lda = 3 * D; g.ldb = D; g.ldc = D;
        g.k_chunk = (int)align_up(ceil_div(M, splits), GBK);
        g.c_split_stride = 3ll * D * D;
        g.drop = drop; g.drop_sid = kDropEmbedding;
        g.drop_on = (news && drop.enabled()) ? 2 : 0;
        NRMS_CHECK_CUDA(launch_gemm_simt(g, false, false, ceil_div(M, g.k_chunk), s, "gemm_wgrad_qkv"));
        rc = reduce_rows(sc.wpart, gv.Wqkv, ceil_div(M, g.k_chunk), 3ll * D * D, 3ll * D * D, 1.f,
                         nullptr, s);
        if (rc) return rc;
    }
    // 6. d_x = d_qkv W_qkv (x the embedding dropout mask for the news encoder)
    if (d_x) {
        GemmArgs g{};
        g.A = sc.d_qkv; g.B = pv.Wqkv; g.C = d_x;
        g.M = M; g.N = D; g.K = 3 * D; g.lda = 3 * D; g.ldb = D; g.ldc = D;
        g.k_chunk = 3 * D;
        g.drop = drop; g.drop_sid = kDropEmbedding;
        g.drop_on = (news && drop.enabled()) ? 3 : 0;
        if (d.gemm_mode == 1) {
            // B(n = d_in, k = qkv column) = W_qkv[k, n]
            NRMS_CHECK_CUDA(tc::pack_b(pv.Wqkv, D, 3 * D, 1, D, sv.pk_qkv_t, s));
            NRMS_CHECK_CUDA(tc::launch(g, sv.pk_qkv_t, 3, s, "gemm_dgrad_qkv"));
        } else {
            NRMS_CHECK_CUDA(launch_gemm_simt(g, true, false, 1, s, "gemm_dgrad_qkv"));
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
                       d_params, d_rows, true, (cudaStream_t)stream);
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
                       d_params, d_x, false, (cudaStream_t)stream);
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
                          nrms_stream_t stream) {
    int rc = score_check(B, C, D);
    if (rc) return rc;
    if (B_global < 1) return fail(NRMS_ERR_BAD_SHAPE, "B_global=%d", B_global);
    NRMS_REQUIRE_PTR(cand); NRMS_REQUIRE_PTR(user); NRMS_REQUIRE_PTR(logits);
    NRMS_REQUIRE_PTR(d_cand); NRMS_REQUIRE_PTR(d_user);
    if (!loss_per_row) return fail(NRMS_ERR_NULL, "loss_per_row is NULL");
    ScoreArgs a{};
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
    NRMS_CHECK_CUDA(cudaMemsetAsync(v.counts, 0, sizeof(int32_t) * vocab, s));
    NRMS_LAUNCH("plan_hist", s, plan_hist_kernel<<<grid_for(n_rows, 256), 256, 0, s>>>(ids, n_rows, vocab, v.counts));
    NRMS_LAUNCH("plan_scan", s, plan_scan_kernel<<<1, 1024, 0, s>>>(v.counts, v.offsets, v.cursor, v.n_valid, vocab));
    NRMS_LAUNCH("plan_fill", s, plan_fill_kernel<<<grid_for(n_rows, 256), 256, 0, s>>>(ids, n_rows, vocab, v.offsets, v.cursor,
                                                          v.perm, v.sorted_id));
    NRMS_LAUNCH("plan_sort_segments", s, plan_sort_segments_kernel<<<grid_for((long long)vocab * 32, 256), 256, 0, s>>>(v.offsets, v.perm,
                                                                                  vocab));
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
    NRMS_LAUNCH("embgrad_reduce", s, embgrad_reduce_kernel<<<grid_for(n_rows, 256, 16), 256, 0, s>>>(v.perm, v.sorted_id, v.offsets,
                                                                   v.n_valid, d_rows, D, d_table));
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

int nrms_dropout_mask(uint64_t seed, uint32_t stream_id, float p, int64_t n, float* out,
                      nrms_stream_t stream) {
    if (n < 1 || p < 0.f || p >= 1.f) return fail(NRMS_ERR_BAD_SHAPE, "n=%lld p=%f", (long long)n, (double)p);
    if (!out) return fail(NRMS_ERR_NULL, "out is NULL");
    NRMS_LAUNCH("dropout_mask", (cudaStream_t)stream, dropout_mask_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(make_dropout(p, seed),
                                                                           stream_id, n, out));
    NRMS_CHECK_CUDA(cudaGetLastError());
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

}  // extern "C"
