// eval.cu — accuracy metrics of top-k lists on the device.
//
// Replaces the per-user Python loops of Evaluator.eval (elliot/evaluation/evaluator.py:117-147) for the
// four metrics the path reports: nDCG (ndcg.py:68-125 with relevance.py:55 discount ln2/ln(r+2) and
// relevance.py:80-82 gains 2^(score-thr+1)-1), HR (hit_rate.py), Precision (precision.py), Recall (recall.py).
// The top-k index tensor produced by the scoring kernels never leaves HBM: each group of lanes owns one user,
// every lane looks one recommended item up in the user's relevant-item row (binary search over the item-sorted
// test CSR) and the group reduces gain*discount and the hit count with a fixed shuffle tree.
// Users without relevant items are skipped, like evaluator.py:121.  Sums are fp64 and deterministic
// (fixed block tree + ordered second pass); the caller divides by the number of evaluated users.
#include "common.cuh"

namespace eb {

constexpr int EVAL_THREADS = 256;
constexpr int EVAL_NOUT = 5;   // n_evaluated, sum nDCG, sum HR, sum Precision, sum Recall

template <int G>
__global__ void __launch_bounds__(EVAL_THREADS) eval_topk_kernel(
    const int32_t *__restrict__ topk, int64_t n_rows, int ld, int k, const int32_t *__restrict__ users,
    const int64_t *__restrict__ indptr, const int32_t *__restrict__ items, const double *__restrict__ gains,
    const double *__restrict__ idcg, const double *__restrict__ disc, double *__restrict__ per_user,
    double *__restrict__ partial) {
    constexpr int UPB = EVAL_THREADS / G;                       // users per block
    __shared__ double acc[EVAL_NOUT][UPB];
    const int g = threadIdx.x / G, lane = threadIdx.x % G;
    const int64_t row = (int64_t)blockIdx.x * UPB + g;
    double v_n = 0, v_ndcg = 0, v_hr = 0, v_p = 0, v_r = 0;
    const bool live = row < n_rows;
    const int64_t u = live ? (users ? (int64_t)users[row] : row) : 0;
    const int64_t lo = live ? indptr[u] : 0, hi = live ? indptr[u + 1] : 0;
    double dcg = 0;
    int hits = 0;
    for (int r = lane; r < k && hi > lo; r += G) {
        const int32_t it = topk[row * ld + r];
        double gain = 0;
        if (it >= 0) {
            int64_t a = lo, b = hi;
            while (a < b) {
                const int64_t m = (a + b) >> 1;
                if (items[m] < it) a = m + 1; else b = m;
            }
            if (a < hi && items[a] == it) gain = gains[a];
        }
        dcg += gain * disc[r];
        hits += gain > 0;
    }
    // all 32 lanes shuffle (groups of one warp may belong to users with and without relevant items)
#pragma unroll
    for (int o = G / 2; o; o >>= 1) {
        dcg += __shfl_xor_sync(0xffffffffu, dcg, o, G);
        hits += __shfl_xor_sync(0xffffffffu, hits, o, G);
    }
    if (hi > lo) {
        v_n = 1;
        v_ndcg = dcg > 0 ? dcg / idcg[u] : 0.0;
        v_hr = hits > 0 ? 1.0 : 0.0;
        v_p = (double)hits / (double)k;
        v_r = (double)hits / (double)(hi - lo);
    }
    if (per_user && live && lane == 0) {
        double *o = per_user + row * 4;
        const double nan = __longlong_as_double(0x7ff8000000000000LL);
        o[0] = v_n ? v_ndcg : nan; o[1] = v_n ? v_hr : nan; o[2] = v_n ? v_p : nan; o[3] = v_n ? v_r : nan;
    }
    if (lane == 0) { acc[0][g] = v_n; acc[1][g] = v_ndcg; acc[2][g] = v_hr; acc[3][g] = v_p; acc[4][g] = v_r; }
    __syncthreads();
    for (int s = UPB / 2; s; s >>= 1) {                         // fixed tree -> deterministic
        if ((int)threadIdx.x < s) {
#pragma unroll
            for (int m = 0; m < EVAL_NOUT; ++m) acc[m][threadIdx.x] += acc[m][threadIdx.x + s];
        }
        __syncthreads();
    }
    if (threadIdx.x < EVAL_NOUT) partial[(int64_t)blockIdx.x * EVAL_NOUT + threadIdx.x] = acc[threadIdx.x][0];
}

// ordered second pass: one warp per output, lanes stride the partials, fixed shuffle tree
__global__ void eval_finish_kernel(const double *__restrict__ partial, int64_t n_blocks, double *__restrict__ out) {
    const int m = threadIdx.x / 32, lane = threadIdx.x % 32;
    double s = 0;
    for (int64_t b = lane; b < n_blocks; b += 32) s += partial[b * EVAL_NOUT + m];
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[m] = s;
}

static int eval_group(int k) { return k <= 8 ? 8 : (k <= 16 ? 16 : 32); }

}  // namespace eb

extern "C" size_t eb_eval_topk_workspace_bytes(int64_t n_rows, int k) {
    const int upb = eb::EVAL_THREADS / eb::eval_group(k);
    return (size_t)((n_rows + upb - 1) / upb) * eb::EVAL_NOUT * sizeof(double) + 64;
}

extern "C" int eb_eval_topk_f64(const int32_t *topk_idx, int64_t n_rows, int ld, int k, const int32_t *users,
                                const int64_t *rel_indptr, const int32_t *rel_items, const double *rel_gains,
                                const double *idcg, const double *discount, double *per_user, double *out,
                                void *workspace, size_t workspace_bytes, void *stream) {
    using namespace eb;
    EB_ARG(out, "null output pointer");
    EB_ARG(n_rows >= 0 && k >= 1 && k <= 1024 && ld >= k, "bad shape (n_rows=%lld k=%d ld=%d)", (long long)n_rows, k, ld);
    cudaStream_t st = (cudaStream_t)stream;
    if (n_rows == 0) {                                          // empty list set: nothing evaluated
        EB_CUDA(cudaMemsetAsync(out, 0, EVAL_NOUT * sizeof(double), st));
        return EB_OK;
    }
    EB_ARG(topk_idx && rel_indptr && rel_items && rel_gains && idcg && discount, "null pointer");
    if (workspace_bytes < eb_eval_topk_workspace_bytes(n_rows, k) || !workspace)
        return set_err(EB_ERR_WORKSPACE, "eval workspace too small: need %zu bytes", eb_eval_topk_workspace_bytes(n_rows, k));
    const int G = eval_group(k);
    const int upb = EVAL_THREADS / G;
    const int64_t blocks = (n_rows + upb - 1) / upb;
    EB_ARG(blocks <= 0x7fffffffLL, "too many rows for one launch");
    double *partial = (double *)workspace;
#define EB_LAUNCH(GV)                                                                                          \
    eval_topk_kernel<GV><<<(unsigned)blocks, EVAL_THREADS, 0, st>>>(topk_idx, n_rows, ld, k, users, rel_indptr, \
                                                                    rel_items, rel_gains, idcg, discount, per_user, partial)
    if (G == 8) EB_LAUNCH(8); else if (G == 16) EB_LAUNCH(16); else EB_LAUNCH(32);
#undef EB_LAUNCH
    EB_CUDA(cudaGetLastError());
    eval_finish_kernel<<<1, 32 * EVAL_NOUT, 0, st>>>(partial, blocks, out);
    EB_CUDA(cudaGetLastError());
    return EB_OK;
}
