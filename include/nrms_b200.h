/*
 * nrms_b200.h — C-ABI of libnrms_b200.so: the B200 (sm_100a) implementation of the NRMS
 * train + scoring hot path of 0215Arthur/Pytorch_News_Recommender.
 *
 * The reference has no FFI: its boundary is a Python nn.Module
 * (MIND_2020/model/nrms_v0.py:218-312) driven by MIND_2020/train_eval.py:156-273.  Each entry
 * point below replaces the ATen op sequence of one reference function; the function it
 * replaces is cited as file:line (paths relative to the reference's MIND_2020/ directory).
 * The Python host module (pytorch_news_recommender_b200/model/nrms_v0.py) binds these with
 * ctypes; INTEGRATION.md shows the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name starts with h_;
 *   - the library never allocates or frees device memory: inputs, outputs, saved
 *     activations and scratch are caller-owned blobs sized by the *_bytes() queries;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*);
 *   - return value 0 = ok, negative = error (see NRMS_ERR_*); nrms_last_error() returns a
 *     thread-local message for the last failing call;
 *   - all matrices are row-major fp32, ids are int64, masks are uint8 (the dtypes the
 *     reference's batches use: data_handler.py:191-234).
 *
 * Flat encoder parameter layout (`params`, `d_params`), one block per encoder, D = d_model,
 * Q = d_query:
 *     [ W_Q (D*D) | W_K (D*D) | W_V (D*D) | b_Q (D) | b_K (D) | b_V (D) |
 *       additive.linear.weight (Q*D) | additive.linear.bias (Q) | attention_query_vector (Q) ]
 * i.e. nn.Linear's [out,in] layout (nrms_v0.py:35-37, 91-93), concatenated.
 */
#ifndef NRMS_B200_H
#define NRMS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NRMS_ABI_VERSION 1

#define NRMS_OK 0
#define NRMS_ERR_BAD_SHAPE (-1)   /* unsupported / inconsistent dims */
#define NRMS_ERR_ALIGN (-2)       /* pointer not 16-byte aligned */
#define NRMS_ERR_CUDA (-3)        /* CUDA launch/runtime error */
#define NRMS_ERR_WORKSPACE (-4)   /* saved/scratch blob too small */
#define NRMS_ERR_NULL (-5)        /* required pointer is NULL */

typedef void* nrms_stream_t; /* cudaStream_t */

/* Shape of one encoder launch.  News encoder: n_seq titles of seq_len words
 * (nrms_v0.py:154-176).  User encoder: n_seq users of seq_len clicked-news vectors
 * (nrms_v0.py:188-199). */
typedef struct nrms_encoder_dims {
    int32_t n_seq;      /* sequences in this launch */
    int32_t seq_len;    /* config.n_words_title or config.history_len */
    int32_t d_model;    /* config.word_embed_size (300) */
    int32_t n_heads;    /* config.num_attention_heads (10) */
    int32_t d_query;    /* config.query_vector_dim (200) */
    int32_t vocab;      /* rows of the word-embedding table (news encoder); 0 for user */
    float dropout_p;    /* config.dropout when training, 0 in eval (nrms_v0.py:137,171-173) */
    int32_t gemm_mode;  /* 0 = fp32 CUDA-core GEMMs, 1 = tcgen05 split-bf16 (bf16x3, fp32-grade),
                           2 = tcgen05 plain bf16 (hi planes only: the "bf16" configs of BASELINE.json,
                           looser parity bound); modes 1/2 need d_model <= 316, d_query <= 208 and
                           an even head dim */
    uint64_t seed;      /* Philox key of this step's dropout masks */
} nrms_encoder_dims;

int nrms_abi_version(void);
const char* nrms_last_error(void);

/* Kernel launches issued by this library since load (bench.py's gpu_launches). */
int64_t nrms_launch_count(void);
/* Opt-in per-kernel timing: CUDA events recorded around every launch on its own stream.
 * nrms_profile_collect synchronises and writes "kernel count total_ms" lines into the HOST
 * buffer h_buf, then clears the log. */
void nrms_profile_enable(int on);
int nrms_profile_collect(char* h_buf, int64_t h_buf_bytes);

/* number of floats in one encoder's flat parameter block */
int64_t nrms_encoder_param_count(int32_t d_model, int32_t d_query);
/* bytes of the activations the forward saves for the backward */
int64_t nrms_encoder_saved_bytes(const nrms_encoder_dims* d);
/* bytes of scratch the backward needs */
int64_t nrms_encoder_scratch_bytes(const nrms_encoder_dims* d);

/* NewsEncoder.forward (nrms_v0.py:154-176): embedding gather (+dropout), 3 projections,
 * scaled-dot-product attention per head (nrms_v0.py:13-23, 46-76), dropout, additive
 * attention pooling (nrms_v0.py:100-126).
 *   ids   [n_seq, seq_len] int64     table [vocab, d_model]     out [n_seq, d_model]
 *   saved: caller-owned blob of nrms_encoder_saved_bytes(d) bytes holding the activations
 *   (needed by the backward; inference callers simply reuse one blob across calls). */
int nrms_news_encoder_fwd(const nrms_encoder_dims* d, const int64_t* ids, const float* table,
                          const float* params, float* out, void* saved, int64_t saved_bytes,
                          nrms_stream_t stream);

/* Autograd mirror of the above (train_eval.py:204).  Writes d_params (overwrite) and the
 * per-token embedding-row gradients d_rows [n_seq*seq_len, d_model] (dropout mask already
 * applied) that nrms_embedding_grad_* scatters into the table gradient. */
int nrms_news_encoder_bwd(const nrms_encoder_dims* d, const int64_t* ids, const float* table,
                          const float* params, const float* d_out, const void* saved,
                          int64_t saved_bytes, void* scratch, int64_t scratch_bytes,
                          float* d_params, float* d_rows, nrms_stream_t stream);

/* The same backward in two calls, for data-parallel training: NRMS_BWD_DATA runs the
 * data-gradient path (d_out -> d_rows, everything the embedding-table gradient needs),
 * NRMS_BWD_PARAMS the weight/bias/query-vector gradients (d_params) from the scratch the first
 * call left.  The table gradient's all-reduce is started between the two and overlaps the
 * second (engine.FusedTrainer).  DATA must precede PARAMS on the same stream and blobs. */
#define NRMS_BWD_DATA 1
#define NRMS_BWD_PARAMS 2
int nrms_news_encoder_bwd_phase(const nrms_encoder_dims* d, const int64_t* ids, const float* table,
                                const float* params, const float* d_out, const void* saved,
                                int64_t saved_bytes, void* scratch, int64_t scratch_bytes,
                                float* d_params, float* d_rows, int32_t phase, nrms_stream_t stream);

/* The user encoder's backward in the same two phases (its weight-gradient half then runs on a side
 * stream underneath the news encoder's data-gradient path; needs a scratch blob of its own). */
int nrms_user_encoder_bwd_phase(const nrms_encoder_dims* d, const float* x, const float* params,
                                const float* d_out, const void* saved, int64_t saved_bytes, void* scratch,
                                int64_t scratch_bytes, float* d_params, float* d_x, int32_t phase,
                                nrms_stream_t stream);

/* UserEncoder.forward (nrms_v0.py:188-199): x [n_seq, seq_len, d_model] -> out [n_seq, d_model] */
int nrms_user_encoder_fwd(const nrms_encoder_dims* d, const float* x, const float* params,
                          float* out, void* saved, int64_t saved_bytes, nrms_stream_t stream);
/* The same encoder with its input rows GATHERED by id from a vector table (cached-vector scoring,
 * BASELINE cfg4: the clicked-news vectors are rows of the news-vector cache built once with
 * get_news_vector, nrms_v0.py:278-299): x[s, l, :] = table[ids[s, l], :], table [d->vocab, d_model].
 * Needs gemm_mode >= 1 (the gather writes the GEMM operand image directly). */
int nrms_user_encoder_fwd_gather(const nrms_encoder_dims* d, const int64_t* ids, const float* table,
                                 const float* params, float* out, void* saved, int64_t saved_bytes,
                                 nrms_stream_t stream);
int nrms_user_encoder_bwd(const nrms_encoder_dims* d, const float* x, const float* params,
                          const float* d_out, const void* saved, int64_t saved_bytes,
                          void* scratch, int64_t scratch_bytes, float* d_params, float* d_x,
                          nrms_stream_t stream);

/* DotProductClickPredictor.forward + candidate masking (nrms_v0.py:205-216, 272-274):
 *   cand [B,C,D], user [B,D], mask [B,C] uint8 (may be NULL) -> logits [B,C] (pad = -1e9) */
int nrms_score_fwd(int32_t B, int32_t C, int32_t D, const float* cand, const float* user,
                   const uint8_t* mask, float* logits, nrms_stream_t stream);
/* get_prediction over cached vectors (nrms_v0.py:301-312 + the -1e9 fill of :272-274):
 *   logits[b, c] = vecs[cand_ids[b, c]] . user[b]; vecs [n_vecs, D], cand_ids [B,S] int64, mask [B,S]
 *   uint8 or NULL.  No [B,S,D] candidate tensor is materialised. */
int nrms_score_cached(int32_t B, int32_t S, int32_t D, const float* vecs, int64_t n_vecs,
                      const int64_t* cand_ids, const float* user, const uint8_t* mask, float* logits,
                      nrms_stream_t stream);
/* its backward: d_logits [B,C] -> d_cand [B,C,D], d_user [B,D] (masked slots get 0) */
int nrms_score_bwd(int32_t B, int32_t C, int32_t D, const float* cand, const float* user,
                   const uint8_t* mask, const float* d_logits, float* d_cand, float* d_user,
                   nrms_stream_t stream);
/* Fused scorer + nn.CrossEntropyLoss vs label 0 (train_eval.py:181,194-195) + its backward.
 * loss_per_row [B] receives logsumexp(s_b)-s_b0.  loss_mean (optional, may be NULL) receives the
 * mean over the B rows of this call — the `loss` of train_eval.py:195 — written by the last CTA
 * to finish, which sums loss_per_row in index order (no floating-point atomics: bitwise
 * repeatable); `ticket` is one uint32 of device memory, zero before the first call and re-armed
 * by the kernel (required when loss_mean is given).  d_cand / d_user are the gradients of the
 * MEAN loss over B_global rows. */
int nrms_score_ce_fwd_bwd(int32_t B, int32_t C, int32_t D, int32_t B_global, const float* cand,
                          const float* user, const uint8_t* mask, float* logits,
                          float* loss_per_row, float* d_cand, float* d_user,
                          float* loss_mean, uint32_t* ticket, nrms_stream_t stream);

/* Deduplicated sparse scatter-add of the embedding-row gradients (the replacement of the 55
 * dense embedding_dense_backward calls).  plan = STABLE radix sort of the
 * (id, row) pairs by vocab row (id 0 = padding_idx, dropped: nrms_v0.py:136), so the rows of
 * one id are summed in row order on every run; runs that cross a 32-row chunk edge go through
 * per-chunk partial slots inside the plan blob and are combined in chunk order (no
 * floating-point atomics: d_table is bitwise repeatable for any id distribution).
 *   plan blob layout is private; size it with nrms_embedding_plan_bytes (~ 120 bytes per row:
 *   sort buffers + 2 partial slots of 384 floats per 32 rows). */
int64_t nrms_embedding_plan_bytes(int64_t n_rows, int32_t vocab);
int nrms_embedding_plan(const int64_t* ids, int64_t n_rows, int32_t vocab, void* plan,
                        int64_t plan_bytes, nrms_stream_t stream);
/* d_table [vocab, D] = sum over rows with the same id (every row of d_table is written) */
int nrms_embedding_grad_dense(const void* plan, int64_t plan_bytes, const float* d_rows,
                              int64_t n_rows, int32_t vocab, int32_t D, float* d_table,
                              nrms_stream_t stream);
/* number of distinct non-pad ids in the plan (device int32 written to *d_unique) */
int nrms_embedding_plan_unique(const void* plan, int64_t plan_bytes, int32_t vocab,
                               int32_t* d_unique, nrms_stream_t stream);

/* torch.optim.Adam defaults (train_eval.py:167,205): betas (0.9,0.999), eps 1e-8, no weight
 * decay, no amsgrad.  step >= 1 is the step being taken.  g may be scaled by grad_scale
 * (1/world_size after a sum-allreduce). */
int nrms_adam_step(float* p, const float* g, float* m, float* v, int64_t n, int32_t step,
                   float lr, float beta1, float beta2, float eps, float grad_scale,
                   nrms_stream_t stream);

/* Ranking metrics of evaluation.py:6-27 for ragged impressions: impression i owns
 * scores[offsets[i] .. offsets[i+1]) and labels likewise (labels 0/1 uint8).
 * max_len >= the longest impression (config.max_candidate_size).
 * out [n_impr, 4] float64 = AUC (roc_auc_score, midrank ties), MRR, nDCG@5, nDCG@10;
 * NaN where the reference yields NaN (single-class impressions). */
int nrms_rank_metrics(const float* scores, const uint8_t* labels, const int64_t* offsets,
                      int64_t n_impr, int32_t max_len, double* out, nrms_stream_t stream);
/* Same, reading padded score rows as train_eval.py:219-227 does: impression i owns
 * scores[i*row_stride .. i*row_stride+len_i), len_i = offsets[i+1]-offsets[i]. */
int nrms_rank_metrics_padded(const float* scores, int64_t row_stride, const uint8_t* labels,
                             const int64_t* offsets, int64_t n_impr, int32_t max_len,
                             double* out, nrms_stream_t stream);
/* Same with BOTH sides padded, the layout an eval batch has (data_handler.py:174-177,236-250):
 * impression i owns scores[i*row_stride ..] and labels[i*label_stride ..], lens[i] real candidates. */
int nrms_rank_metrics_rows(const float* scores, int64_t row_stride, const uint8_t* labels,
                           int64_t label_stride, const int64_t* lens, int64_t n_impr, int32_t max_len,
                           double* out, nrms_stream_t stream);

/* Batch assembly on the device: MyDataset.__getitem__ + default_collate of data_handler.py:185-250
 * for the keys the NRMS path reads, in ONE launch.  Inputs are the sample matrices packed once on
 * the host (front-aligned, zero padded: browsed_ids [N,H], candidate_ids [N,S], their lengths) and
 * the title table [n_news, T] (news id = row + 1; id 0 -> all-zero title), all int64 in HBM;
 * index [B] picks the samples.  Outputs: browsed_ids [B,H], browsed_lens [B], browsed_titles
 * [B,H,T], browsed_mask [B,H] u8, candidate_ids [B,S], candidate_titles [B,S,T], candidate_mask
 * [B,S] u8. */
int nrms_assemble_batch(const int64_t* index, int32_t B, const int64_t* browsed_ids,
                        const int64_t* browsed_lens, const int64_t* candidate_ids,
                        const int64_t* candidate_lens, const int64_t* titles, int64_t n_news,
                        int32_t H, int32_t S, int32_t T, int64_t* o_browsed_ids,
                        int64_t* o_browsed_lens, int64_t* o_browsed_titles, uint8_t* o_browsed_mask,
                        int64_t* o_candidate_ids, int64_t* o_candidate_titles,
                        uint8_t* o_candidate_mask, nrms_stream_t stream);

/* Rank lists of the submission writer (train_eval.py:279-285 `_cal_test`, written to
 * sumbit_*.txt by train_eval.py:335-339): ranks[i, j] = 1 + position of candidate j in
 * argsort(-scores[i, :lens[i]]) for j < lens[i] (ties: lower index first), 0 for j >= lens[i].
 * scores / ranks are padded rows of row_stride elements; lens int64 [n_impr], clipped to
 * [0, row_stride]. */
int nrms_rank_positions(const float* scores, int64_t row_stride, const int64_t* lens,
                        int64_t n_impr, int32_t* ranks, nrms_stream_t stream);

/* Gather rows: out[i,:] = src[idx[i]-base,:] (idx < base -> zeros).  Used to build
 * [B,H,D]/[B,C,D] from the cached news-vector table keyed by browsed_ids/candidate_ids
 * (news row + 1, 0 = pad: data_handler.py:88,100) and, with int64 title tables, for the
 * on-device batch assembly of data_handler.py:206-228. */
int nrms_gather_rows_f32(const float* src, int64_t n_src, int32_t D, const int64_t* idx,
                         int64_t n_idx, int64_t base, float* out, nrms_stream_t stream);
int nrms_gather_rows_i64(const int64_t* src, int64_t n_src, int32_t D, const int64_t* idx,
                         int64_t n_idx, int64_t base, int64_t* out, nrms_stream_t stream);

/* Test hook: writes the dropout multiplier (0 or 1/(1-p)) that the encoder kernels apply to
 * element (r, c) of a [n_rows, n_cols] activation in stream `stream_id` (1 = embedding dropout
 * nrms_v0.py:137 over the gathered rows [n_titles*T, D], 2 = context dropout nrms_v0.py:171-173
 * over the attention output [n_titles*T, D]).  out is [n_rows, n_cols] fp32. */
int nrms_dropout_mask(uint64_t seed, uint32_t stream_id, float p, int64_t n_rows, int32_t n_cols,
                      float* out, nrms_stream_t stream);

/* Self-test of the tcgen05 GEMM layer (gemm_img.cuh): packs fp32 row-major operands into
 * split-bf16 images and runs one of the three operand orientations the path uses:
 *   variant 0: C[M,N] = A[M,K] * B[N,K]^T   forward projection        (K-major  x K-major)
 *   variant 1: C[M,N] = A[M,K] * B[K,N]     data gradient, N <= 320   (K-major  x MN-major)
 *   variant 2: C[M,N] = A[K,M]^T * B[K,N]   weight gradient, N <= 320 (MN-major x MN-major, split-K)
 *   variant 3: variant 0 on CTA pairs (tcgen05.mma.cta_group::2, M = 256 per pair of CTAs)
 * work: caller-owned scratch of nrms_gemm_selftest_bytes() bytes.  N % 4 == 0. */
int64_t nrms_gemm_selftest_bytes(int32_t variant, int32_t M, int32_t N, int32_t K);
int nrms_gemm_selftest(int32_t variant, const float* A, const float* B, float* C, int32_t M,
                       int32_t N, int32_t K, void* work, int64_t work_bytes, nrms_stream_t stream);

/* out-of-range id check: *d_flag |= 1 if any id < 0 or >= vocab */
int nrms_validate_ids(const int64_t* ids, int64_t n, int64_t vocab, int32_t* d_flag,
                      nrms_stream_t stream);

/* ---- the `nrms` sibling variant (reference model/nrms.py) ---------------------
 * Its sub-modules as separate entry points (the host side composes them: pytorch_news_recommender_b200/
 * model/nrms.py), all fp32 in / fp32 out, row-major, caller-owned work blobs.
 *
 * nn.Linear (nrms.py:65-66 the Q/K/V and output projections, :91 the additive projection, :226-230
 * news_dense): y[M,N] = x[M,K] W[N,K]^T + bias[N] on the tcgen05 image GEMMs (fp32-grade bf16x3);
 * backward: dx[M,K] = dy W (dx may be NULL), dW[N,K] = dy^T x, dbias[N] = column sums of dy (may be NULL;
 * it comes out of the weight-gradient GEMM through a ones column appended to x).
 * N % 4 == 0, K % 4 == 0; work >= nrms_linear_work_bytes(M, N, K) serves both directions. */
int64_t nrms_linear_work_bytes(int32_t M, int32_t N, int32_t K);
int nrms_linear_fwd(const float* x, const float* W, const float* bias, float* y, int32_t M, int32_t N,
                    int32_t K, void* work, int64_t work_bytes, nrms_stream_t stream);
int nrms_linear_bwd(const float* x, const float* W, const float* dy, float* dx, float* dW, float* dbias,
                    int32_t M, int32_t N, int32_t K, void* work, int64_t work_bytes, nrms_stream_t stream);

/* nn.Dropout on a [n_rows, n_cols] activation (nrms.py:254): y = x * m with m the multiplier
 * nrms_dropout_mask() reports for (seed, stream_id, p); the backward is the same call on dy.
 * stream_id 3 = candidate vectors, 4 = history vectors (5 = attention probabilities, applied inside
 * nrms_masked_attention_*). */
int nrms_dropout_apply(uint64_t seed, uint32_t stream_id, float p, int64_t n_rows, int32_t n_cols,
                       const float* x, float* y, nrms_stream_t stream);

/* Attention.forward (nrms.py:26-49) for all heads: qkv [B, L, 3*heads*dk] (Q | K | V side by side),
 * mask [B, L] uint8 (1 = real slot; NULL = no mask): score(i, j) = -1e9 unless slots i AND j are real
 * (:38-41); probs [B, heads, L, L] = softmax BEFORE dropout (saved for the backward); dropout of rate
 * p_drop on the probabilities (:45-47; multiplier of element (b, h, i, j) = nrms_dropout_mask(seed, 5, p)
 * at row (b*heads + h)*L + i, column j); ctx [B, L, heads*dk].  L <= 128, dk <= 128 and the item's
 * tiles must fit shared memory (4 L ld + 2 L ceil8(L) + 64 L floats <= 227 KB, ld ~ dk + 4: the
 * reference's 50-slot history with 64-wide heads takes 90 KB).
 * Backward: d_qkv [B, L, 3*heads*dk] from d_ctx; masked scores pass no gradient. */
int nrms_masked_attention_fwd(const float* qkv, const uint8_t* mask, int32_t B, int32_t L, int32_t heads,
                              int32_t dk, float p_drop, uint64_t seed, float* probs, float* ctx,
                              nrms_stream_t stream);
int nrms_masked_attention_bwd(const float* qkv, const uint8_t* mask, const float* probs, const float* d_ctx,
                              int32_t B, int32_t L, int32_t heads, int32_t dk, float p_drop, uint64_t seed,
                              float* d_qkv, nrms_stream_t stream);

/* AdditiveAttention.forward (nrms.py:98-117) after its Linear: t [B, L, Q] = x W^T + b (pre-tanh),
 * qv [Q], x [B, L, E], mask [B, L] or NULL (padded slots scored -1e9, :112-113) ->
 * alpha [B, L] softmax weights, out [B, E] = sum_i alpha_i x_i.
 * Backward: d_t [B, L, Q], d_x [B, L, E] (the pooling's own share alpha_i * d_out; the Linear's share
 * comes from nrms_linear_bwd), d_qv [Q]; work: B*Q floats. */
int nrms_masked_pool_fwd(const float* t, const float* qv, const float* x, const uint8_t* mask, int32_t B,
                         int32_t L, int32_t Q, int32_t E, float* alpha, float* out, nrms_stream_t stream);
int nrms_masked_pool_bwd(const float* t, const float* qv, const float* x, const uint8_t* mask,
                         const float* alpha, const float* d_out, int32_t B, int32_t L, int32_t Q, int32_t E,
                         float* d_t, float* d_x, float* d_qv, float* work, nrms_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* NRMS_B200_H */
