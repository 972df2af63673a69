// score_topk.cu — exact full-catalogue scoring + train mask + per-user top-k on CUDA cores.
//
// Replaces MFModel.get_user_predictions (BPRMF_model.py:70-85) and
// BPRMF_batch_model.predict/get_top_k (BPRMF_batch_model.py:82-88) in the table's own
// precision (fp32 or fp64).  It is the exact path for small catalogues and the
// re-check path for users the tensor-core kernel (score_topk_tc.cu) cannot certify.
// One CTA per user at a time: warps stream item rows (coalesced), the score row lives in
// an L2-resident scratch line, masking is a CSR scatter of -inf, selection is k rounds of
// a block-wide (value desc, index asc) arg-max.
#include <math_constants.h>

#include "common.cuh"

namespace eb {

template <typename T>
struct ScoreParams {
    const T *U, *V, *bias;
    int32_t n_items;
    int d, ld;
    const int64_t *mask_indptr;
    const int32_t *mask_indices;
    const int32_t *users;
    const int32_t *positions;  // optional: row q works on user_begin+positions[q] and writes output row positions[q]
    const int32_t *n_rows_dev; // optional: number of rows actually present (device side), <= n_sel
    int32_t user_begin;
    int64_t n_sel;
    int k;
    int32_t *out_idx;
    T *out_val;
    T *scratch;  // gridDim.x rows of n_items
};

template <typename T> __device__ __forceinline__ T neg_inf();
template <> __device__ __forceinline__ float neg_inf<float>() { return -CUDART_INF_F; }
template <> __device__ __forceinline__ double neg_inf<double>() { return -CUDART_INF; }

template <typename T>
__device__ __forceinline__ bool better(T v, int i, T bv, int bi) {
    return v > bv || (v == bv && v > neg_inf<T>() && i < bi);
}

template <typename T>
__global__ void __launch_bounds__(256) score_topk_exact_kernel(const ScoreParams<T> p) {
    extern __shared__ unsigned char smem_raw[];
    T *su = reinterpret_cast<T *>(smem_raw);  // user row, d entries
    __shared__ T red_v[8];
    __shared__ int red_i[8];
    __shared__ int win_i;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T *s = p.scratch + (int64_t)blockIdx.x * p.n_items;
    const int64_t n_rows = p.n_rows_dev ? min((int64_t)*p.n_rows_dev, p.n_sel) : p.n_sel;
    for (int64_t q = blockIdx.x; q < n_rows; q += gridDim.x) {
        const int64_t qo = p.positions ? (int64_t)p.positions[q] : q;
        const int u = p.users ? p.users[q] : p.user_begin + (int)qo;
        __syncthreads();
        for (int k = threadIdx.x; k < p.d; k += blockDim.x) su[k] = p.U[(int64_t)u * p.ld + k];
        __syncthreads();
        // phase 1: scores
        for (int it = warp; it < p.n_items; it += 8) {
            const T *vr = p.V + (int64_t)it * p.ld;
            T acc = 0;
            for (int k = lane; k < p.d; k += 32) acc = fma(su[k], vr[k], acc);   // score_topk_tc.cu re-rank mirrors this order
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
            if (lane == 0) s[it] = (p.bias ? p.bias[it] : (T)0) + acc;
        }
        __syncthreads();
        // phase 2: train items -> -inf  (BPRMF_model.py:73-74, BPRMF_batch_model.py:88)
        if (p.mask_indptr) {
            const int64_t beg = p.mask_indptr[u], end = p.mask_indptr[u + 1];
            for (int64_t m = beg + threadIdx.x; m < end; m += blockDim.x) s[p.mask_indices[m]] = neg_inf<T>();
        }
        __syncthreads();
        // phase 3: k rounds of arg-max, ties -> lower index
        for (int r = 0; r < p.k; r++) {
            T bv = neg_inf<T>();
            int bi = 0x7fffffff;
            for (int it = threadIdx.x; it < p.n_items; it += blockDim.x) {
                const T v = s[it];
                if (better<T>(v, it, bv, bi)) { bv = v; bi = it; }
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const T ov = __shfl_xor_sync(0xffffffffu, bv, off);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
                if (better<T>(ov, oi, bv, bi)) { bv = ov; bi = oi; }
            }
            if (lane == 0) { red_v[warp] = bv; red_i[warp] = bi; }
            __syncthreads();
            if (warp == 0) {
                bv = lane < 8 ? red_v[lane] : neg_inf<T>();
                bi = lane < 8 ? red_i[lane] : 0x7fffffff;
#pragma unroll
                for (int off = 4; off > 0; off >>= 1) {
                    const T ov = __shfl_xor_sync(0xffffffffu, bv, off);
                    const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
                    if (better<T>(ov, oi, bv, bi)) { bv = ov; bi = oi; }
                }
                if (lane == 0) {
                    const bool ok = bv > neg_inf<T>();
                    p.out_idx[qo * p.k + r] = ok ? bi : -1;
                    p.out_val[qo * p.k + r] = bv;
                    win_i = ok ? bi : -1;
                    if (ok) s[bi] = neg_inf<T>();
                }
            }
            __syncthreads();
            if (win_i < 0) {  // nothing finite left: pad the tail
                for (int rr = r + 1 + threadIdx.x; rr < p.k; rr += blockDim.x) {
                    p.out_idx[qo * p.k + rr] = -1;
                    p.out_val[qo * p.k + rr] = neg_inf<T>();
                }
                break;
            }
        }
    }
}

static int64_t score_ctas(int64_t n_sel) {
    int64_t c = (int64_t)sm_count() * 4;
    return n_sel < c ? (n_sel < 1 ? 1 : n_sel) : c;
}

template <typename T>
static int score_topk_exact(const T *U, const T *V, const T *bias, int32_t n_items, int d, int ld,
                            const int64_t *mask_indptr, const int32_t *mask_indices, const int32_t *users,
                            int32_t user_begin, int64_t n_sel, int k, int32_t *out_idx, T *out_val, void *workspace,
                            size_t workspace_bytes, void *stream, const int32_t *positions = nullptr,
                            const int32_t *n_rows_dev = nullptr) {
    EB_ARG(U && V && out_idx && out_val, "null pointer");
    EB_ARG(d >= 1 && ld >= d && n_items >= 1 && k >= 1, "bad shape d=%d ld=%d n_items=%d k=%d", d, ld, n_items, k);
    EB_ARG((mask_indptr == nullptr) == (mask_indices == nullptr), "mask CSR: both or neither");
    if (n_sel <= 0) return EB_OK;
    int64_t ctas = score_ctas(n_sel);
    const size_t row = sizeof(T) * (size_t)n_items;
    if (workspace_bytes < row || !workspace)
        return set_err(EB_ERR_WORKSPACE, "workspace %zu < one score row %zu", workspace_bytes, row);
    if ((size_t)ctas * row > workspace_bytes) ctas = (int64_t)(workspace_bytes / row);
    ScoreParams<T> p{U, V, bias, n_items, d, ld, mask_indptr, mask_indices, users, positions, n_rows_dev, user_begin, n_sel, k, out_idx,
                     out_val, (T *)workspace};
    score_topk_exact_kernel<T><<<(unsigned)ctas, 256, sizeof(T) * (size_t)d, (cudaStream_t)stream>>>(p);
    EB_CUDA(cudaGetLastError());
    return EB_OK;
}

}  // namespace eb

extern "C" size_t eb_score_topk_workspace_bytes(int64_t n_sel, int32_t n_items, int elem_size) {
    return (size_t)eb::score_ctas(n_sel) * (size_t)n_items * (size_t)elem_size;
}

extern "C" int eb_score_topk_f32(const float *U, const float *V, const float *item_bias, int32_t n_items, int d, int ld,
                                 const int64_t *mask_indptr, const int32_t *mask_indices, const int32_t *users,
                                 int32_t user_begin, int64_t n_sel, int k, int32_t *out_idx, float *out_val,
                                 void *workspace, size_t workspace_bytes, void *stream) {
    return eb::score_topk_exact<float>(U, V, item_bias, n_items, d, ld, mask_indptr, mask_indices, users, user_begin,
                                       n_sel, k, out_idx, out_val, workspace, workspace_bytes, stream);
}

extern "C" int eb_score_topk_f64(const double *U, const double *V, const double *item_bias, int32_t n_items, int d,
                                 int ld, const int64_t *mask_indptr, const int32_t *mask_indices, const int32_t *users,
                                 int32_t user_begin, int64_t n_sel, int k, int32_t *out_idx, double *out_val,
                                 void *workspace, size_t workspace_bytes, void *stream) {
    return eb::score_topk_exact<double>(U, V, item_bias, n_items, d, ld, mask_indptr, mask_indices, users, user_begin,
                                        n_sel, k, out_idx, out_val, workspace, workspace_bytes, stream);
}

// re-check entry used by score_topk_tc.cu: row q scores user user_begin+positions[q] and writes output row positions[q]
extern "C" int eb_score_topk_f32_mapped(const float *U, const float *V, const float *item_bias, int32_t n_items, int d,
                                        int ld, const int64_t *mask_indptr, const int32_t *mask_indices,
                                        const int32_t *positions, int32_t user_begin, int64_t n_sel, int k,
                                        int32_t *out_idx, float *out_val, void *workspace, size_t workspace_bytes,
                                        void *stream) {
    return eb::score_topk_exact<float>(U, V, item_bias, n_items, d, ld, mask_indptr, mask_indices, nullptr, user_begin,
                                       n_sel, k, out_idx, out_val, workspace, workspace_bytes, stream, positions);
}

extern "C" int eb_score_topk_f32_mapped_dev(const float *U, const float *V, const float *item_bias, int32_t n_items, int d,
                                            int ld, const int64_t *mask_indptr, const int32_t *mask_indices,
                                            const int32_t *positions, const int32_t *n_rows_dev, int32_t user_begin,
                                            int64_t n_sel_max, int k, int32_t *out_idx, float *out_val, void *workspace,
                                            size_t workspace_bytes, void *stream) {
    EB_ARG(positions && n_rows_dev, "null pointer");
    return eb::score_topk_exact<float>(U, V, item_bias, n_items, d, ld, mask_indptr, mask_indices, nullptr, user_begin,
                                       n_sel_max, k, out_idx, out_val, workspace, workspace_bytes, stream, positions, n_rows_dev);
}
