// metrics.cuh — per-impression ranking metrics on device.
// Reference: evaluation.py:6-27 (dcg_score / ndcg_score / mrr_score / auc_score) called per
// impression from train_eval.py:219-227 on rank_score[i][:len(y_true[i])].
//   order = argsort(score)[::-1]; ties follow a stable ascending sort reversed, i.e. the
//   HIGHER index ranks first (what numpy does for n <= 16 and for tie-free input).
//   AUC = sklearn roc_auc_score = (#(pos,neg) pairs with s_pos > s_neg + 0.5 * #ties) / (P*N)
// One warp per impression, scores/labels staged in shared memory, O(n^2) compares
// (n <= 300 candidate slots, mean ~37).
#pragma once
#include "common.cuh"

namespace nrms {

constexpr int kMetricWarps = 8;

__global__ void __launch_bounds__(kMetricWarps * 32) rank_metrics_kernel(
    const float* __restrict__ scores, long long row_stride /* <0: ragged by offsets */,
    const uint8_t* __restrict__ labels, const int64_t* __restrict__ offsets, long long n_impr,
    int max_n, double* __restrict__ out, long long label_stride = -1 /* >= 0: labels padded [n_impr, label_stride] */,
    const int64_t* __restrict__ lens = nullptr /* with label_stride: candidates per impression */) {
    extern __shared__ __align__(16) unsigned char smraw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* s = reinterpret_cast<float*>(smraw) + (size_t)warp * max_n;
    uint8_t* y = smraw + (size_t)kMetricWarps * max_n * sizeof(float) + (size_t)warp * max_n;
    const double kNaN = __longlong_as_double(0x7ff8000000000000LL);
    for (long long imp = (long long)blockIdx.x * kMetricWarps + warp; imp < n_impr;
         imp += (long long)gridDim.x * kMetricWarps) {
        const long long o0 = label_stride >= 0 ? imp * label_stride : offsets[imp];
        const int n = (int)(label_stride >= 0 ? lens[imp] : offsets[imp + 1] - o0);
        double* res = out + imp * 4;
        if (n <= 0 || n > max_n) {
            if (lane < 4) res[lane] = kNaN;
            continue;
        }
        const float* sp = row_stride < 0 ? scores + o0 : scores + imp * row_stride;
        int pos = 0;
        for (int j = lane; j < n; j += 32) {
            s[j] = sp[j];
            const uint8_t yy = labels[o0 + j];
            y[j] = yy;
            pos += yy ? 1 : 0;
        }
        pos = __reduce_add_sync(0xffffffffu, pos);
        const int neg = n - pos;
        __syncwarp();
        double auc_num = 0.0, mrr = 0.0, dcg5 = 0.0, dcg10 = 0.0;
        for (int i = lane; i < n; i += 32) {
            if (!y[i]) continue;
            const float si = s[i];
            int gt = 0, eq_after = 0, less_neg = 0, eq_neg = 0;
            for (int j = 0; j < n; ++j) {
                const float sj = s[j];
                const bool isneg = y[j] == 0;
                gt += sj > si;
                eq_after += (sj == si) && (j > i);
                less_neg += isneg && (sj < si);
                eq_neg += isneg && (sj == si);
            }
            const int rank = gt + eq_after;  // 0-based position in argsort(score)[::-1]
            auc_num += (double)less_neg + 0.5 * (double)eq_neg;
            mrr += 1.0 / (double)(rank + 1);
            const double disc = 1.0 / log2((double)(rank + 2));
            if (rank < 5) dcg5 += disc;
            if (rank < 10) dcg10 += disc;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            auc_num += __shfl_xor_sync(0xffffffffu, auc_num, o);
            mrr += __shfl_xor_sync(0xffffffffu, mrr, o);
            dcg5 += __shfl_xor_sync(0xffffffffu, dcg5, o);
            dcg10 += __shfl_xor_sync(0xffffffffu, dcg10, o);
        }
        if (lane == 0) {
            double idcg5 = 0.0, idcg10 = 0.0;
            for (int r = 0; r < min(pos, 10); ++r) {
                const double d = 1.0 / log2((double)(r + 2));
                if (r < 5) idcg5 += d;
                idcg10 += d;
            }
            res[0] = (pos > 0 && neg > 0) ? auc_num / ((double)pos * (double)neg) : kNaN;
            res[1] = pos > 0 ? mrr / (double)pos : kNaN;
            res[2] = pos > 0 ? dcg5 / idcg5 : kNaN;
            res[3] = pos > 0 ? dcg10 / idcg10 : kNaN;
        }
        __syncwarp();
    }
}

// Rank list of the submission writer (train_eval.py:279-285, `_cal_test`): for impression i with
// n = lens[i] real candidates, ranks[i, j] = 1 + position of candidate j in argsort(-score[:n]).
// Ties keep the lower index first (a stable sort; numpy's default sort leaves tie order
// unspecified).  Slots j >= n get rank 0.  One warp per impression, O(n^2) compares.
__global__ void __launch_bounds__(kMetricWarps * 32) rank_positions_kernel(
    const float* __restrict__ scores, long long row_stride, const int64_t* __restrict__ lens,
    long long n_impr, int max_n, int32_t* __restrict__ ranks) {
    extern __shared__ __align__(16) unsigned char smraw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* s = reinterpret_cast<float*>(smraw) + (size_t)warp * max_n;
    for (long long imp = (long long)blockIdx.x * kMetricWarps + warp; imp < n_impr;
         imp += (long long)gridDim.x * kMetricWarps) {
        long long n64 = lens[imp];
        const int n = (int)(n64 < 0 ? 0 : (n64 > max_n ? max_n : n64));
        const float* row = scores + imp * row_stride;
        int32_t* out = ranks + imp * row_stride;
        for (int j = lane; j < n; j += 32) s[j] = row[j];
        __syncwarp();
        for (int j = lane; j < (int)row_stride; j += 32) {
            int r = 0;
            if (j < n) {
                const float sj = s[j];
                int before = 0;
                for (int k = 0; k < n; ++k) {
                    const float sk = s[k];
                    before += (sk > sj) || (sk == sj && k < j);
                }
                r = before + 1;
            }
            out[j] = r;
        }
        __syncwarp();
    }
}

}  // namespace nrms
