// gmf.cu — Generalized Matrix Factorization (NeuMF's MF branch as a model of its own).
// Reference: elliot/recommender/neural/GeneralizedMF/generalized_matrix_factorization_model.py:18-92 (is_edge_weight_train =
// True, the default): out = sigmoid((U[u] * I[i]) . h), Keras BinaryCrossentropy (batch mean, probabilities clipped to
// [1e-7, 1-1e-7]), Adam on U, I and the edge weights h; sampler dataset/samplers/pointwise_pos_neg_sampler.py:24-48.
//
//   eb_gmf_step_grads     ONE fused kernel per batch: gather the two rows -> product -> dot with h -> sigmoid -> loss ->
//                         gradient -> vector-atomic scatter into the dense gradient tables (+ dh, loss)
//   eb_gmf_scale_rows     U' = U * h (per column): scoring is then a plain U' . I^T, ranked by the logit (sigmoid is
//                         monotone), so get_recs/get_top_k (:76-92) run on the tensor-core scoring kernel
//   eb_sigmoid_inplace    the k kept logits -> probabilities
// The pointwise sampler lives in bpr_train.cu (eb_pointwise_sample_philox: it shares the BPR sampler's draws).
// TensorFlow parity is UNPINNED (oracle/tf_models.py::gmf_forward_backward is the checker).
#include <math_constants.h>

#include "common.cuh"

namespace eb {

// one warp per sample, lanes over f/4 float4 (f % 4 == 0, f <= 128: one float4 per lane)
__global__ void __launch_bounds__(256) gmf_step_kernel(const float *__restrict__ U, const float *__restrict__ I, int64_t ld, int f,
                                                       const float *__restrict__ h, const int32_t *__restrict__ u,
                                                       const int32_t *__restrict__ it, const float *__restrict__ label, int64_t n,
                                                       float invn, float *dU, float *dI, float *dh, double *loss) {
    const int lane = threadIdx.x & 31;
    int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int c = lane * 4;
    const bool on = c < f;
    const float4 hv = on ? *reinterpret_cast<const float4 *>(h + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 acc_h = make_float4(0.f, 0.f, 0.f, 0.f);
    float acc_loss = 0.f;
    for (; w < n; w += nw) {
        const int uu = u[w], ii = it[w];
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
        if (on) {
            a = *reinterpret_cast<const float4 *>(U + (int64_t)uu * ld + c);
            b = *reinterpret_cast<const float4 *>(I + (int64_t)ii * ld + c);
        }
        const float4 pm = make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w);
        float part = pm.x * hv.x + pm.y * hv.y + pm.z * hv.z + pm.w * hv.w;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
        const float pr = 1.f / (1.f + __expf(-part));
        const float y = label[w];
        const float pc = fminf(fmaxf(pr, 1e-7f), 1.f - 1e-7f);        // Keras backend.binary_crossentropy clipping
        if (lane == 0) acc_loss += -(y * __logf(pc) + (1.f - y) * __logf(1.f - pc));
        const float dl = (pr > 1e-7f && pr < 1.f - 1e-7f) ? (pr - y) * invn : 0.f;
        if (on && dl != 0.f) {
            acc_h.x += dl * pm.x; acc_h.y += dl * pm.y; acc_h.z += dl * pm.z; acc_h.w += dl * pm.w;
            const float4 g = make_float4(dl * hv.x, dl * hv.y, dl * hv.z, dl * hv.w);           // d loss / d pm
            red_add_v4(dU + (int64_t)uu * ld + c, make_float4(g.x * b.x, g.y * b.y, g.z * b.z, g.w * b.w));
            red_add_v4(dI + (int64_t)ii * ld + c, make_float4(g.x * a.x, g.y * a.y, g.z * a.z, g.w * a.w));
        }
    }
    if (on && (acc_h.x != 0.f || acc_h.y != 0.f || acc_h.z != 0.f || acc_h.w != 0.f)) red_add_v4(dh + c, acc_h);
    if (loss && lane == 0 && acc_loss != 0.f) atomicAdd(loss, (double)acc_loss * (double)invn);
}

__global__ void __launch_bounds__(256) scale_rows_kernel(const float *__restrict__ src, int64_t ld, int64_t rows, int f,
                                                         const float *__restrict__ h, float *__restrict__ dst, int64_t ldd) {
    const int64_t total = rows * (f / 4);
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = e / (f / 4); const int c = (int)(e - r * (f / 4)) * 4;
        const float4 a = *reinterpret_cast<const float4 *>(src + r * ld + c), w = *reinterpret_cast<const float4 *>(h + c);
        *reinterpret_cast<float4 *>(dst + r * ldd + c) = make_float4(a.x * w.x, a.y * w.y, a.z * w.z, a.w * w.w);
    }
}

__global__ void __launch_bounds__(256) sigmoid_kernel(float *x, int64_t n) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x)
        x[e] = 1.f / (1.f + __expf(-x[e]));                          // -inf (padding) -> 0
}

static inline unsigned ggrid(int64_t threads) {
    int64_t g = (threads + 255) / 256; const int64_t cap = (int64_t)sm_count() * 8;
    return (unsigned)(g > cap ? cap : (g < 1 ? 1 : g));
}

}  // namespace eb

using namespace eb;

extern "C" int eb_gmf_step_grads(const float *U, const float *I, int64_t ld, int f, const float *h, const int32_t *u, const int32_t *it,
                                 const float *label, int64_t n, int64_t mean_over, float *dU, float *dI, float *dh, double *loss,
                                 void *stream) {
    EB_ARG(U && I && h && u && it && label && dU && dI && dh, "null pointer");
    EB_ARG(f >= 4 && f % 4 == 0 && f <= 128 && ld >= f && ld % 4 == 0 && mean_over >= 1 && n >= 0, "f must be a multiple of 4, <= 128");
    if (n == 0) return EB_OK;
    gmf_step_kernel<<<ggrid(n * 32), 256, 0, (cudaStream_t)stream>>>(U, I, ld, f, h, u, it, label, n, 1.f / (float)mean_over, dU, dI, dh, loss);
    EB_CUDA(cudaGetLastError());
    return EB_OK;
}

extern "C" int eb_gmf_scale_rows(const float *src, int64_t ld, int64_t rows, int f, const float *h, float *dst, int64_t ldd, void *stream) {
    EB_ARG(src && h && dst && f >= 4 && f % 4 == 0 && ld >= f && ldd >= f && ld % 4 == 0 && ldd % 4 == 0 && rows >= 0, "bad argument");
    if (rows == 0) return EB_OK;
    scale_rows_kernel<<<ggrid(rows * (f / 4)), 256, 0, (cudaStream_t)stream>>>(src, ld, rows, f, h, dst, ldd);
    EB_CUDA(cudaGetLastError());
    return EB_OK;
}

extern "C" int eb_sigmoid_inplace(float *x, int64_t n, void *stream) {
    EB_ARG(x && n >= 0, "bad argument");
    if (n == 0) return EB_OK;
    sigmoid_kernel<<<ggrid(n), 256, 0, (cudaStream_t)stream>>>(x, n);
    EB_CUDA(cudaGetLastError());
    return EB_OK;
}
