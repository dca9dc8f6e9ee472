// train_kernels.cuh — click scorer + cross-entropy, embedding-gradient dedupe/scatter, Adam,
// row gathers.  All HBM-bound; grids are sized against the 148 SMs.
#pragma once
#include "common.cuh"

namespace nrms {

// ---------------------------------------------------------------------------------------
// DotProductClickPredictor.forward (nrms_v0.py:205-216) + masked_fill (nrms_v0.py:272-274)
// [+ nn.CrossEntropyLoss vs label 0 (train_eval.py:181,194-195) and the backward of both].
// One CTA per impression; one warp per candidate slot (looping).
// ---------------------------------------------------------------------------------------
struct ScoreArgs {
    const float* cand;   // [B, C, D]
    const float* user;   // [B, D]
    const uint8_t* mask; // [B, C] or nullptr
    float* logits;       // [B, C]
    const float* d_logits; // [B, C] (score_bwd only)
    float* loss_rows;    // [B]    (fused only)
    float* loss_mean;    // [1]    (fused, optional): mean of loss_rows, summed in index order by the last CTA
    unsigned int* ticket;  // [1]  (fused, with loss_mean): zero before the first call; the kernel resets it
    float* d_cand;       // [B, C, D]
    float* d_user;       // [B, D]
    int B, C, D;
    float inv_batch;     // 1 / B_global
};

// MODE 0: logits only. MODE 1: fused logits + CE + grads. MODE 2: grads from d_logits.
template <int MODE>
__global__ void __launch_bounds__(256) score_kernel(const ScoreArgs a) {
    extern __shared__ float sm[];  // C floats: logits then d_logits
    const int b = blockIdx.x, C = a.C, D = a.D;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const float* u = a.user + (long long)b * D;
    const float* cb = a.cand + (long long)b * C * D;
    if (MODE != 2) {
        for (int c = warp; c < C; c += nw) {
            const float* cr = cb + (long long)c * D;
            float s = 0.f;
            for (int d = lane; d < D; d += 32) s = fmaf(__ldg(cr + d), __ldg(u + d), s);
            s = warp_sum(s);
            if (lane == 0) {
                if (a.mask && a.mask[(long long)b * C + c] == 0) s = -1e9f;
                sm[c] = s;
                a.logits[(long long)b * C + c] = s;
            }
        }
        __syncthreads();
    }
    if (MODE == 0) return;
    if (MODE == 1) {
        if (warp == 0) {
            float mx = -INFINITY;
            for (int c = lane; c < C; c += 32) mx = fmaxf(mx, sm[c]);
            mx = warp_max(mx);
            float den = 0.f;
            for (int c = lane; c < C; c += 32) den += expf(sm[c] - mx);
            den = warp_sum(den);
            const float lse = mx + logf(den);
            const float s0 = sm[0];
            __syncwarp();
            for (int c = lane; c < C; c += 32) {
                float g = expf(sm[c] - lse) - (c == 0 ? 1.f : 0.f);
                if (a.mask && a.mask[(long long)b * C + c] == 0) g = 0.f;  // masked_fill grad
                sm[c] = g * a.inv_batch;
            }
            if (lane == 0) a.loss_rows[b] = lse - s0;
        }
    } else {
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            float g = a.d_logits[(long long)b * C + c];
            if (a.mask && a.mask[(long long)b * C + c] == 0) g = 0.f;
            sm[c] = g;
        }
    }
    __syncthreads();
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        const float ud = __ldg(u + d);
        float du = 0.f;
        for (int c = 0; c < C; ++c) {
            const float g = sm[c];
            du = fmaf(g, __ldg(cb + (long long)c * D + d), du);
            a.d_cand[((long long)b * C + c) * D + d] = g * ud;
        }
        a.d_user[(long long)b * D + d] = du;
    }
    if (MODE == 1 && a.loss_mean != nullptr) {
        // mean over the B rows without a second launch and without floating-point atomics: the CTA that
        // draws the last ticket sums loss_rows in index order (bitwise repeatable) and re-arms the ticket
        __shared__ unsigned int s_last;
        __shared__ float s_part[8];
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            s_last = atomicAdd(a.ticket, 1u) == (unsigned)a.B - 1u ? 1u : 0u;
        }
        __syncthreads();
        if (s_last) {
            __threadfence();
            // fixed assignment of rows to lanes and a fixed shuffle tree: the same bits every run
            float acc = 0.f;
            for (int i = threadIdx.x; i < a.B; i += blockDim.x) acc += __ldcg(a.loss_rows + i);
            acc = warp_sum(acc);
            if (lane == 0) s_part[warp] = acc;
            __syncthreads();
            if (threadIdx.x == 0) {
                float t = 0.f;
                for (int w = 0; w < nw; ++w) t += s_part[w];
                *a.loss_mean = t / (float)a.B;
                *a.ticket = 0u;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// Cached-vector scoring (BASELINE cfg4; reference hooks get_prediction nrms_v0.py:301-312 over
// get_news_vector :278-289): logits[b, c] = vecs[cand_ids[b, c]] . user[b] straight from the news-vector
// cache — no [B, S, D] candidate tensor is materialised, padded slots (mask == 0) get -1e9
// (nrms_v0.py:272-274) without touching the cache.  One CTA per impression, one warp per candidate
// slot (looping); the cache (65 k x 1.2 KB) is L2-resident.
// ---------------------------------------------------------------------------------------
struct ScoreCachedArgs {
    const float* vecs;        // [n_vecs, D]
    long long n_vecs;
    const int64_t* cand_ids;  // [B, S]
    const float* user;        // [B, D]
    const uint8_t* mask;      // [B, S] or nullptr
    float* logits;            // [B, S]
    int B, S, D;
};
__global__ void __launch_bounds__(256) score_cached_kernel(const ScoreCachedArgs a) {
    extern __shared__ float su[];   // D floats: this impression's user vector
    const int b = blockIdx.x, S = a.S, D = a.D;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int d = threadIdx.x; d < D; d += blockDim.x) su[d] = __ldg(a.user + (long long)b * D + d);
    __syncthreads();
    const int d4 = D >> 2;          // D % 4 == 0
    for (int c = warp; c < S; c += nw) {
        const long long o = (long long)b * S + c;
        float s = -1e9f;
        if (a.mask == nullptr || a.mask[o] != 0) {
            const long long id = a.cand_ids[o];
            s = 0.f;
            if (id >= 0 && id < a.n_vecs) {
                const float4* row = reinterpret_cast<const float4*>(a.vecs + id * D);
                for (int i = lane; i < d4; i += 32) {
                    const float4 v = __ldg(row + i);
                    const float4 u = *reinterpret_cast<const float4*>(su + 4 * i);
                    s = fmaf(v.x, u.x, s); s = fmaf(v.y, u.y, s); s = fmaf(v.z, u.z, s); s = fmaf(v.w, u.w, s);
                }
            }
            s = warp_sum(s);
        }
        if (lane == 0) a.logits[o] = s;
    }
}

// ---------------------------------------------------------------------------------------
// Embedding gradient: stable sort of the token rows by vocab id, then a load-balanced segmented
// reduction (one warp per 32 sorted rows) into the dense table gradient.
// Replaces 55 x (zero-fill [V,D] + scatter + accumulate) of the reference.
//
// Everything here is DETERMINISTIC, also for words that occur thousands of times in a step (Zipf / real
// text): the rows of a word are summed in ascending row order, in fixed 32-row pieces, and pieces are
// combined in piece order — no floating-point atomics and no order handed out by an atomic cursor.
//   1. counts / offsets: histogram (integer atomics: exact) + two-level exclusive scan;
//   2. (id, row) pairs sorted by id with a stable LSD radix sort (rows of one id stay in row order);
//   3. one warp per 32 sorted rows sums runs of equal id; a run that lies inside its chunk is stored, a
//      run that crosses a chunk edge leaves its piece in a partial slot of that chunk;
//   4. one warp per multi-chunk word adds that word's pieces in chunk order and stores the row.
// plan blob: int32 counts[V] | offsets[V+1] | perm[n] | sorted_id[n] | n_valid | block_tot[ceil(V/1024)]
//            | tmp_key[n] | tmp_val[n] | radix hist[512 * ceil(n/2048)] | its scan [.. + 1] | its block totals
//            | float part[2][ceil(n/32)][kPartLd]
// ---------------------------------------------------------------------------------------
constexpr int kPartLd = 384;          // floats per partial slot (the kernels support D <= 384)
constexpr int kRsBlock = 2048;        // keys per CTA of the radix passes (8 warps x 8 x 32)
constexpr int kRsMaxBins = 512;

struct PlanView {
    int32_t* counts;
    int32_t* offsets;
    int32_t* perm;
    int32_t* sorted_id;
    int32_t* n_valid;
    int32_t* block_tot;
    int32_t* tmp_key;
    int32_t* tmp_val;
    int32_t* rhist;       // [bins][nblk] digit counts of every 2048-key block (digit-major)
    int32_t* rscan;       // exclusive scan of rhist (+ the total)
    int32_t* rblock_tot;  // totals of the scan's 1024-entry blocks
    float* part;          // [2][n_chunks][kPartLd]: slot 0 = piece of a run that began in an EARLIER chunk,
                          //                           slot 1 = piece of a run that continues into the NEXT chunk
    long long n_chunks;
};
inline int64_t plan_ints(int64_t n_rows, int32_t vocab) {
    const int64_t nh = (int64_t)kRsMaxBins * ceil_div64(n_rows, kRsBlock);
    return 2ll * vocab + 1 + 2 * n_rows + 4 + (vocab + 1023) / 1024 + 2 * n_rows + nh + (nh + 1) + ceil_div64(nh, 1024);
}
inline int64_t plan_bytes(int64_t n_rows, int32_t vocab) {
    return align_up((int64_t)sizeof(int32_t) * plan_ints(n_rows, vocab), 256) +
           (int64_t)sizeof(float) * 2 * ceil_div64(n_rows, 32) * kPartLd;
}
inline PlanView plan_view(void* blob, int64_t n_rows, int32_t vocab) {
    PlanView v;
    int32_t* p = reinterpret_cast<int32_t*>(blob);
    v.counts = p;
    v.offsets = v.counts + vocab;
    v.perm = v.offsets + vocab + 1;
    v.sorted_id = v.perm + n_rows;
    v.n_valid = v.sorted_id + n_rows;
    v.block_tot = v.n_valid + 4;
    v.tmp_key = v.block_tot + (vocab + 1023) / 1024;
    v.tmp_val = v.tmp_key + n_rows;
    v.rhist = v.tmp_val + n_rows;
    const int64_t nh = (int64_t)kRsMaxBins * ceil_div64(n_rows, kRsBlock);
    v.rscan = v.rhist + nh;
    v.rblock_tot = v.rscan + nh + 1;
    v.part = reinterpret_cast<float*>(reinterpret_cast<char*>(blob) +
                                      align_up((int64_t)sizeof(int32_t) * plan_ints(n_rows, vocab), 256));
    v.n_chunks = ceil_div64(n_rows, 32);
    return v;
}

__global__ void plan_hist_kernel(const int64_t* __restrict__ ids, long long n, int vocab,
                                 int32_t* __restrict__ counts) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const long long id = ids[i];
        if (id > 0 && id < vocab) atomicAdd(counts + id, 1);
    }
}

// Exclusive scan of counts[V] -> offsets[V+1] in two small launches of ceil(V/1024) CTAs:
// (1) per-CTA totals, (2) every CTA sums the totals of the CTAs before it (at most a few hundred
// values) and scans its own 1024 counts.  n_valid = total.
constexpr int kScanBlock = 1024;

__device__ __forceinline__ int32_t block_inclusive_scan(int32_t x, int32_t* warp_tot /*[32]*/, int32_t* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int32_t inc = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int32_t n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += n;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const int32_t w = warp_tot[lane];
        int32_t winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int32_t n = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += n;
        }
        warp_tot[lane] = winc - w;  // exclusive warp offsets
        if (lane == 31) *total = winc;
    }
    __syncthreads();
    return inc + warp_tot[warp];
}

__global__ void __launch_bounds__(kScanBlock) plan_block_totals_kernel(const int32_t* __restrict__ counts,
                                                                      int32_t* __restrict__ block_tot, int vocab) {
    __shared__ int32_t warp_tot[32];
    __shared__ int32_t total;
    const int i = blockIdx.x * kScanBlock + threadIdx.x;
    block_inclusive_scan(i < vocab ? counts[i] : 0, warp_tot, &total);
    if (threadIdx.x == 0) block_tot[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kScanBlock) plan_scan_kernel(const int32_t* __restrict__ counts,
                                                              const int32_t* __restrict__ block_tot,
                                                              int32_t* __restrict__ offsets,
                                                              int32_t* __restrict__ n_valid, int vocab) {
    __shared__ int32_t warp_tot[32];
    __shared__ int32_t total;
    __shared__ int32_t base_s;
    // sum of the totals of the CTAs before this one
    int32_t part = 0;
    for (int b = threadIdx.x; b < (int)blockIdx.x; b += kScanBlock) part += block_tot[b];
    block_inclusive_scan(part, warp_tot, &total);
    if (threadIdx.x == 0) base_s = total;
    __syncthreads();
    const int32_t base = base_s;
    __syncthreads();
    const int i = blockIdx.x * kScanBlock + threadIdx.x;
    const int32_t c = i < vocab ? counts[i] : 0;
    const int32_t inc = block_inclusive_scan(c, warp_tot, &total);
    if (i < vocab) offsets[i] = base + inc - c;
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) {
        offsets[vocab] = base + total;
        if (n_valid) *n_valid = base + total;
    }
}

// ---- stable LSD radix sort of (key = vocab id, value = token row) --------------------------------
// key of a padding / out-of-range id = vocab (sorts behind every real id); values start as 0..n-1, so after
// the stable passes the rows of one id are in ascending row order.
__global__ void rsort_init_kernel(const int64_t* __restrict__ ids, long long n, int vocab,
                                  int32_t* __restrict__ key, int32_t* __restrict__ val) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const long long id = ids[i];
        key[i] = (id > 0 && id < vocab) ? (int32_t)id : vocab;
        val[i] = (int32_t)i;
    }
}
// digit histogram of each 2048-key block: hist[digit * nblk + block]
__global__ void __launch_bounds__(256) rsort_hist_kernel(const int32_t* __restrict__ key, long long n, int shift,
                                                        int bins, int32_t* __restrict__ hist, int nblk) {
    __shared__ int32_t h[kRsMaxBins];
    for (int i = threadIdx.x; i < bins; i += blockDim.x) h[i] = 0;
    __syncthreads();
    const long long base = (long long)blockIdx.x * kRsBlock;
    for (int i = threadIdx.x; i < kRsBlock; i += blockDim.x)
        if (base + i < n) atomicAdd(&h[(key[base + i] >> shift) & (bins - 1)], 1);
    __syncthreads();
    for (int i = threadIdx.x; i < bins; i += blockDim.x) hist[(long long)i * nblk + blockIdx.x] = h[i];
}
// (the exclusive scan of hist[bins * nblk], digit-major — all blocks of digit 0, then digit 1, ... — is the
// two-level scan above: plan_block_totals_kernel + plan_scan_kernel over 1024-entry blocks)
// stable scatter: warp w of a block owns keys [256 w, 256 w + 256) of the block and walks them in order, 32
// at a time; the rank of a key among the block's earlier keys of the same digit = (count in earlier warps) +
// (count in this warp's earlier groups) + (earlier lanes of its group with the same digit)
__global__ void __launch_bounds__(256) rsort_scatter_kernel(const int32_t* __restrict__ key, const int32_t* __restrict__ val,
                                                           long long n, int shift, int bins,
                                                           const int32_t* __restrict__ base, int nblk,
                                                           int32_t* __restrict__ key_out, int32_t* __restrict__ val_out) {
    __shared__ int32_t wh[8][kRsMaxBins];      // per-warp digit counts, then exclusive prefix over the warps
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 8 * kRsMaxBins; i += blockDim.x) (&wh[0][0])[i] = 0;
    __syncthreads();
    const long long b0 = (long long)blockIdx.x * kRsBlock + warp * 256;
    int32_t k[8], v[8], rank[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const long long i = b0 + j * 32 + lane;
        const bool ok = i < n;
        k[j] = ok ? key[i] : 0;
        v[j] = ok ? val[i] : 0;
        const int d = ok ? (k[j] >> shift) & (bins - 1) : -1;
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        const int before = __popc(peers & ((1u << lane) - 1u));
        rank[j] = ok ? wh[warp][d] + before : 0;
        __syncwarp();
        if (ok && before == 0) wh[warp][d] += __popc(peers);      // the group's first lane of this digit
        __syncwarp();
    }
    __syncthreads();
    for (int d = threadIdx.x; d < bins; d += blockDim.x) {         // exclusive prefix over the 8 warps, per digit
        int32_t run = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            const int32_t c = wh[w][d];
            wh[w][d] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const long long i = b0 + j * 32 + lane;
        if (i < n) {
            const int d = (k[j] >> shift) & (bins - 1);
            const int32_t pos = base[(long long)d * nblk + blockIdx.x] + wh[warp][d] + rank[j];
            key_out[pos] = k[j];
            val_out[pos] = v[j];
        }
    }
}

// zero-fill (float4) — the dense gradient is written in full every step
__global__ void zero_kernel(float4* __restrict__ p, long long n4) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
         i += (long long)gridDim.x * blockDim.x)
        p[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// One warp per chunk of 32 sorted rows.  Runs of equal id are summed in registers (ascending row order).
// A run that lies inside the chunk AND is its word's whole segment is stored into the table gradient; a run
// that crosses a chunk edge leaves its piece in the chunk's partial slot (0: the run began in an earlier
// chunk; 1: it continues into the next one) for embgrad_fixup_kernel.  D % 4 == 0, D <= 384.
__global__ void __launch_bounds__(256) embgrad_reduce_kernel(
    const int32_t* __restrict__ perm, const int32_t* __restrict__ sorted_id,
    const int32_t* __restrict__ offsets, const int32_t* __restrict__ n_valid_p,
    const float* __restrict__ d_rows, int D, float* __restrict__ d_table, float* __restrict__ part,
    long long n_chunks) {
    const int lane = threadIdx.x & 31;
    const long long warp_g = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int n_valid = *n_valid_p;
    const int d4 = D >> 2;
    for (long long chunk = warp_g; chunk * 32 < n_valid; chunk += nwarps) {
        const int base = (int)(chunk * 32);
        const int cnt = min(32, n_valid - base);
        const int my_row = lane < cnt ? perm[base + lane] : 0;
        const int my_id = lane < cnt ? sorted_id[base + lane] : -1;
        float4 acc[3];
        int cur = __shfl_sync(0xffffffffu, my_id, 0);
        acc[0] = acc[1] = acc[2] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = 0; r <= cnt; ++r) {
            const int id = r < cnt ? __shfl_sync(0xffffffffu, my_id, r) : -2;
            if (id != cur) {
                // flush run [.., r) of id `cur`
                const int seg_beg = offsets[cur], seg_end = offsets[cur + 1];
                float* dst;
                if (seg_beg >= base && seg_end <= base + cnt) dst = d_table + (long long)cur * D;     // whole segment
                else dst = part + ((seg_beg < base ? 0ll : n_chunks) + chunk) * kPartLd;              // a piece
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const int col4 = lane + 32 * c;
                    if (col4 < d4) reinterpret_cast<float4*>(dst)[col4] = acc[c];
                    acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
                cur = id;
            }
            if (r < cnt) {
                const int row = __shfl_sync(0xffffffffu, my_row, r);
                const float4* src = reinterpret_cast<const float4*>(d_rows + (long long)row * D);
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const int col4 = lane + 32 * c;
                    if (col4 < d4) {
                        const float4 v = __ldg(src + col4);
                        acc[c].x += v.x; acc[c].y += v.y; acc[c].z += v.z; acc[c].w += v.w;
                    }
                }
            }
        }
    }
}
// One warp per word whose segment spans several chunks: its pieces — the head piece in slot 1 of the first
// chunk, then slot 0 of every following chunk — are added in chunk order (four loads in flight, adds in order).
__global__ void __launch_bounds__(256) embgrad_fixup_kernel(const int32_t* __restrict__ offsets, int vocab, int D,
                                                           const float* __restrict__ part, long long n_chunks,
                                                           float* __restrict__ d_table) {
    const int lane = threadIdx.x & 31;
    const long long warp_g = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int d4 = D >> 2;
    for (long long v = warp_g; v < vocab; v += nwarps) {
        const int beg = offsets[v], end = offsets[v + 1];
        if (end <= beg) continue;
        const int c0 = beg >> 5, c1 = (end - 1) >> 5;
        if (c0 == c1) continue;                           // stored by the reduction kernel
        float4 acc[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int col4 = lane + 32 * c;
            acc[c] = col4 < d4 ? __ldg(reinterpret_cast<const float4*>(part + (n_chunks + c0) * kPartLd) + col4)
                               : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        for (int ch = c0 + 1; ch <= c1; ch += 4) {
            float4 t[4][3];
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const int col4 = lane + 32 * c;
                    t[u][c] = (ch + u <= c1 && col4 < d4)
                                  ? __ldg(reinterpret_cast<const float4*>(part + (long long)(ch + u) * kPartLd) + col4)
                                  : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (ch + u <= c1) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        acc[c].x += t[u][c].x; acc[c].y += t[u][c].y; acc[c].z += t[u][c].z; acc[c].w += t[u][c].w;
                    }
                }
        }
        float* dst = d_table + v * D;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int col4 = lane + 32 * c;
            if (col4 < d4) reinterpret_cast<float4*>(dst)[col4] = acc[c];
        }
    }
}

__global__ void plan_unique_kernel(const int32_t* __restrict__ counts, int vocab,
                                   int32_t* __restrict__ out) {
    int local = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < vocab; i += gridDim.x * blockDim.x)
        local += counts[i] > 0 ? 1 : 0;
    local = __reduce_add_sync(0xffffffffu, local);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(out, local);
}

// ---------------------------------------------------------------------------------------
// torch.optim.Adam (defaults) — train_eval.py:167,205.  Mirrors torch's single-tensor
// formulation: m = lerp(m, g, 1-b1); v = b2 v + (1-b2) g^2;
// p -= (lr/bc1) * m / (sqrt(v)/sqrt(bc2) + eps)
// ---------------------------------------------------------------------------------------
struct AdamArgs {
    float* p;
    const float* g;
    float* m;
    float* v;
    long long n;
    float beta1, beta2, eps, step_size, bc2_sqrt, grad_scale;
};

__device__ __forceinline__ void adam1(float& p, float g, float& m, float& v, const AdamArgs& a) {
    g *= a.grad_scale;
    m = m + (g - m) * (1.f - a.beta1);
    v = a.beta2 * v + (1.f - a.beta2) * g * g;
    const float denom = sqrtf(v) / a.bc2_sqrt + a.eps;
    p = p - a.step_size * (m / denom);
}

__global__ void __launch_bounds__(256) adam_kernel(const AdamArgs a) {
    const long long n4 = a.n >> 2;
    float4* p4 = reinterpret_cast<float4*>(a.p);
    const float4* g4 = reinterpret_cast<const float4*>(a.g);
    float4* m4 = reinterpret_cast<float4*>(a.m);
    float4* v4 = reinterpret_cast<float4*>(a.v);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
         i += (long long)gridDim.x * blockDim.x) {
        float4 p = p4[i], m = m4[i], v = v4[i];
        const float4 g = g4[i];
        adam1(p.x, g.x, m.x, v.x, a);
        adam1(p.y, g.y, m.y, v.y, a);
        adam1(p.z, g.z, m.z, v.z, a);
        adam1(p.w, g.w, m.w, v.w, a);
        p4[i] = p; m4[i] = m; v4[i] = v;
    }
    // tail (n % 4)
    const long long t = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < a.n) adam1(a.p[t], a.g[t], a.m[t], a.v[t], a);
}

// ---------------------------------------------------------------------------------------
// row gathers (cached news vectors by news id; title tokens by news id)
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void gather_rows_kernel(const T* __restrict__ src, long long n_src, int D,
                                   const int64_t* __restrict__ idx, long long n_idx,
                                   long long base, T* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const long long warp_g = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long r = warp_g; r < n_idx; r += nwarps) {
        const long long s = idx[r] - base;
        const bool ok = s >= 0 && s < n_src;
        for (int d = lane; d < D; d += 32)
            out[r * D + d] = ok ? __ldg(src + s * D + d) : (T)0;
    }
}

// test hook: the multiplier (0 or 1/(1-p)) the encoder kernels apply to element (r, c)
__global__ void dropout_mask_kernel(Dropout dr, uint32_t sid, long long n_rows, int n_cols, float* out) {
    const int groups = ceil_div(n_cols, 8);
    const long long total = n_rows * groups;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / groups;
        const int g = (int)(i - r * groups);
        const uint32_t keep = dr.enabled() ? dr.keep8(sid, (uint64_t)r, (uint32_t)g) : 0xffu;
        for (int j = 0; j < 8 && g * 8 + j < n_cols; ++j)
            out[r * n_cols + g * 8 + j] = ((keep >> j) & 1u) ? dr.scale : 0.f;
    }
}

__global__ void validate_ids_kernel(const int64_t* ids, long long n, long long vocab,
                                    int32_t* flag) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x)
        if (ids[i] < 0 || ids[i] >= vocab) atomicOr(flag, 1);
}

// ---- batch assembly (MyDataset.__getitem__ + default_collate, data_handler.py:185-250) ----------
// One launch builds a whole batch from device-resident sample matrices: a warp owns one
// (sample, slot) of the H history + S candidate slots, copies the news id, the mask byte and the
// title row of that news (id = title row + 1; id 0 or out of range -> the all-zero title).
struct AssembleArgs {
    const int64_t* index;            // [B] sample numbers
    const int64_t* browsed_ids;      // [N, H]
    const int64_t* browsed_lens;     // [N]
    const int64_t* candidate_ids;    // [N, S]
    const int64_t* candidate_lens;   // [N]
    const int64_t* titles;           // [n_news, T]
    long long n_news;
    int B, H, S, T;
    int64_t* o_browsed_ids;          // [B, H]
    int64_t* o_browsed_lens;         // [B]
    int64_t* o_browsed_titles;       // [B, H, T]
    uint8_t* o_browsed_mask;         // [B, H]
    int64_t* o_candidate_ids;        // [B, S]
    int64_t* o_candidate_titles;     // [B, S, T]
    uint8_t* o_candidate_mask;       // [B, S]
};
__global__ void __launch_bounds__(256) assemble_batch_kernel(const AssembleArgs a) {
    const int lane = threadIdx.x & 31;
    const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int N = a.H + a.S;
    for (long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < (long long)a.B * N; w += warps) {
        const int b = (int)(w / N), slot = (int)(w - (long long)b * N);
        const long long src = a.index[b];
        const bool hist = slot < a.H;
        const int j = hist ? slot : slot - a.H;
        const long long id = hist ? a.browsed_ids[src * a.H + j] : a.candidate_ids[src * a.S + j];
        const long long len = hist ? a.browsed_lens[src] : a.candidate_lens[src];
        const long long o = hist ? (long long)b * a.H + j : (long long)b * a.S + j;
        if (lane == 0) {
            (hist ? a.o_browsed_ids : a.o_candidate_ids)[o] = id;
            (hist ? a.o_browsed_mask : a.o_candidate_mask)[o] = (uint8_t)(j < len);
            if (slot == 0) a.o_browsed_lens[b] = len;
        }
        int64_t* dst = (hist ? a.o_browsed_titles : a.o_candidate_titles) + o * a.T;
        const bool real = id >= 1 && id <= a.n_news;
        const int64_t* row = a.titles + (real ? id - 1 : 0) * a.T;
        for (int t = lane; t < a.T; t += 32) dst[t] = real ? __ldg(row + t) : 0;
    }
}

}  // namespace nrms
