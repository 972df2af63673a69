// neumf.cu — the non-GEMM pieces of NeuMF (elliot/recommender/neural/NeuMF/neural_matrix_factorization_model.py).
//
//   eb_neumf_gather     x0 = [U_mlp[u] | I_mlp[i]], p = U_mf[u] * I_mf[i]                      (:74-84)
//   eb_neumf_head       logit = b + w . [p | h3]; prob = sigmoid; Keras BinaryCrossentropy (mean, probs clipped to
//                       [1e-7, 1-1e-7]); d_logit; dP, dH3 (with the last ReLU's mask), dw, db   (:84-86,95-106)
//   eb_relu_bwd         d_pre = d_out * (out > 0)
//   eb_neumf_scatter    embedding gradients into dense gradient tables (vector atomics)
//   eb_neumf_sample     pointwise sampler: every train pair with label 1 plus m uniform non-train items with
//                       label 0 (NeuMF/custom_sampler.py:27-48 distribution; Philox stream, no set-dedup of negatives)
//   eb_neumf_pair_h1    get_recs (:119-144): first MLP layer for ALL (user, item) pairs of a user block from the
//                       factorised pre-activations A_u[u] + A_i[i] + b1, written as the bf16 operand of layer 2
//   eb_neumf_pair_head  final probabilities for the block (mf part + last layer)
// Dense layers run through eb_gemm_bf16_tn.  TensorFlow parity is UNPINNED (oracle/tf_models.py restatement).
#include <cuda_bf16.h>
#include <math_constants.h>

#include "common.cuh"

namespace eb {

__device__ __forceinline__ void nred4(float *p, float4 v) {
    asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// one warp per sample, lanes over f/4 float4 (f % 4 == 0, f <= 128)
__global__ void __launch_bounds__(256) neumf_gather_kernel(const float *Umf, const float *Imf, const float *Umlp, const float *Imlp,
                                                           int f, int64_t ldt, const int32_t *u, const int32_t *it, int64_t n,
                                                           float *x0, int64_t ldx, float *pm, int64_t ldp) {
    const int lane = threadIdx.x & 31;
    int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (; w < n; w += nw) {
        const int uu = u[w], ii = it[w];
        for (int c = lane * 4; c < f; c += 128) {
            const float4 a = *reinterpret_cast<const float4 *>(Umlp + (int64_t)uu * ldt + c);
            const float4 b = *reinterpret_cast<const float4 *>(Imlp + (int64_t)ii * ldt + c);
            *reinterpret_cast<float4 *>(x0 + w * ldx + c) = a;
            *reinterpret_cast<float4 *>(x0 + w * ldx + f + c) = b;
            const float4 m1 = *reinterpret_cast<const float4 *>(Umf + (int64_t)uu * ldt + c);
            const float4 m2 = *reinterpret_cast<const float4 *>(Imf + (int64_t)ii * ldt + c);
            *reinterpret_cast<float4 *>(pm + w * ldp + c) = make_float4(m1.x * m2.x, m1.y * m2.y, m1.z * m2.z, m1.w * m2.w);
        }
    }
}

// one warp per sample; wp = [w_mf (f) | w_mlp (f)], h3 is the last hidden layer AFTER relu
__global__ void __launch_bounds__(256) neumf_head_kernel(const float *pm, int64_t ldp, const float *h3, int64_t ldh, int f,
                                                         const float *wp, const float *bp, const float *label, int64_t n,
                                                         float *dpm, float *dh3, float *dwp, float *dbp, double *loss,
                                                         float *prob_out, float invn) {
    const int lane = threadIdx.x & 31;
    int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    float acc_w[8];                                  // this lane's columns c = lane + 32 j (2f <= 256)
#pragma unroll
    for (int j = 0; j < 8; j++) acc_w[j] = 0.f;
    float acc_b = 0.f, acc_loss = 0.f;
    for (; w < n; w += nw) {
        float part = 0.f;
        for (int c = lane; c < 2 * f; c += 32) part += wp[c] * (c < f ? pm[w * ldp + c] : h3[w * ldh + c - f]);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
        const float logit = part + bp[0];
        const float pr = 1.f / (1.f + __expf(-logit));
        if (prob_out) { if (lane == 0) prob_out[w] = pr; }
        if (!label) continue;
        const float y = label[w];
        const float pc = fminf(fmaxf(pr, 1e-7f), 1.f - 1e-7f);        // Keras backend.binary_crossentropy clipping
        const bool open = pr > 1e-7f && pr < 1.f - 1e-7f;
        if (lane == 0) acc_loss += -(y * __logf(pc) + (1.f - y) * __logf(1.f - pc));
        const float dl = open ? (pr - y) * invn : 0.f;                // d mean-BCE / d logit
        if (lane == 0) acc_b += dl;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int c = lane + 32 * j;
            if (c < 2 * f) {
                const float feat = c < f ? pm[w * ldp + c] : h3[w * ldh + c - f];
                acc_w[j] += dl * feat;
                if (c < f) dpm[w * ldp + c] = dl * wp[c];
                else dh3[w * ldh + c - f] = feat > 0.f ? dl * wp[c] : 0.f;      // ReLU mask of the last hidden layer
            }
        }
    }
    if (!label) return;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const int c = lane + 32 * j;
        if (c < 2 * f && acc_w[j] != 0.f) atomicAdd(dwp + c, acc_w[j]);
    }
    if (lane == 0) {
        if (acc_b != 0.f) atomicAdd(dbp, acc_b);
        if (loss && acc_loss != 0.f) atomicAdd(loss, (double)acc_loss * (double)invn);
    }
}

__global__ void __launch_bounds__(256) relu_bwd_kernel(const float *dout, const float *out, float *dpre, int64_t n, __nv_bfloat16 *copy) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        const float d = out[e] > 0.f ? dout[e] : 0.f;
        dpre[e] = d;
        if (copy) copy[e] = __float2bfloat16_rn(d);          // the operand copy the next two GEMMs read (row length % 8 == 0: same layout)
    }
}

__global__ void __launch_bounds__(256) neumf_scatter_kernel(const float *Umf, const float *Imf, int f, int64_t ldt, const int32_t *u,
                                                            const int32_t *it, int64_t n, const float *dpm, int64_t ldp,
                                                            const float *dx0, int64_t ldx, float *dUmf, float *dImf, float *dUmlp,
                                                            float *dImlp) {
    const int lane = threadIdx.x & 31;
    int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (; w < n; w += nw) {
        const int uu = u[w], ii = it[w];
        for (int c = lane * 4; c < f; c += 128) {
            const float4 g = *reinterpret_cast<const float4 *>(dpm + w * ldp + c);
            const float4 a = *reinterpret_cast<const float4 *>(Umf + (int64_t)uu * ldt + c);
            const float4 b = *reinterpret_cast<const float4 *>(Imf + (int64_t)ii * ldt + c);
            nred4(dUmf + (int64_t)uu * ldt + c, make_float4(g.x * b.x, g.y * b.y, g.z * b.z, g.w * b.w));
            nred4(dImf + (int64_t)ii * ldt + c, make_float4(g.x * a.x, g.y * a.y, g.z * a.z, g.w * a.w));
            nred4(dUmlp + (int64_t)uu * ldt + c, *reinterpret_cast<const float4 *>(dx0 + w * ldx + c));
            nred4(dImlp + (int64_t)ii * ldt + c, *reinterpret_cast<const float4 *>(dx0 + w * ldx + f + c));
        }
    }
}

// sample s in [0, nnz*(1+m)): positive index p = s / (1+m), slot = s % (1+m)
__global__ void __launch_bounds__(256) neumf_sample_kernel(int32_t n_users, int32_t n_items, const int64_t *indptr,
                                                           const int32_t *indices, int m, uint64_t seed, int64_t total,
                                                           int32_t *ou, int32_t *oi, float *ol) {
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < total; s += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = s / (1 + m);
        const int slot = (int)(s - p * (1 + m));
        int lo = 0, hi = n_users;                    // user owning CSR position p: last u with indptr[u] <= p
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (indptr[mid] <= p) lo = mid; else hi = mid; }
        const int u = lo;
        ou[s] = u;
        if (slot == 0) { oi[s] = indices[p]; ol[s] = 1.f; continue; }
        const int64_t beg = indptr[u]; const int len = (int)(indptr[u + 1] - beg);
        uint32_t r[4]; uint32_t attempt = 0; int q = 4; int cand;
        do {
            if (q == 4) { Philox::gen(seed, (uint64_t)s, attempt++, r); q = 0; }
            cand = (int)bounded(r[q++], (uint32_t)n_items);
        } while (contains_sorted(indices + beg, len, cand) && attempt < 1024u);
        oi[s] = cand; ol[s] = 0.f;
    }
}

// H1[(ub, i)] = relu(Au[ub] + Ai[i] + b1) as bf16, row = ub * n_items + i, width h1 (multiple of 8)
__global__ void __launch_bounds__(256) neumf_pair_h1_kernel(const float *Au, int64_t ldau, const float *Ai, int64_t ldai, const float *b1,
                                                            int n_ub, int n_items, int h1, __nv_bfloat16 *out, int64_t ldo) {
    const int64_t total = (int64_t)n_ub * n_items * (h1 / 4);
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int c4 = (int)(e % (h1 / 4));
        const int64_t pair = e / (h1 / 4);
        const int it = (int)(pair % n_items), ub = (int)(pair / n_items);
        const float4 a = *reinterpret_cast<const float4 *>(Au + (int64_t)ub * ldau + c4 * 4);
        const float4 b = *reinterpret_cast<const float4 *>(Ai + (int64_t)it * ldai + c4 * 4);
        const float4 c = *reinterpret_cast<const float4 *>(b1 + c4 * 4);
        __nv_bfloat162 lo = __floats2bfloat162_rn(fmaxf(a.x + b.x + c.x, 0.f), fmaxf(a.y + b.y + c.y, 0.f));
        __nv_bfloat162 hi = __floats2bfloat162_rn(fmaxf(a.z + b.z + c.z, 0.f), fmaxf(a.w + b.w + c.w, 0.f));
        __nv_bfloat162 *o = reinterpret_cast<__nv_bfloat162 *>(out + pair * ldo + c4 * 4);
        o[0] = lo; o[1] = hi;
    }
}

// the same in fp32 (checking mode, ops.exact_gemm): H1 stays unrounded
__global__ void __launch_bounds__(256) neumf_pair_h1_f32_kernel(const float *Au, int64_t ldau, const float *Ai, int64_t ldai, const float *b1,
                                                                int n_ub, int n_items, int h1, float *out, int64_t ldo) {
    const int64_t total = (int64_t)n_ub * n_items * h1;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(e % h1);
        const int64_t pair = e / h1;
        const int it = (int)(pair % n_items), ub = (int)(pair / n_items);
        out[pair * ldo + c] = fmaxf(Au[(int64_t)ub * ldau + c] + Ai[(int64_t)it * ldai + c] + b1[c], 0.f);
    }
}

// prob[(ub, i)] = sigmoid(bp + w_mf . (U_mf[u0+ub] * I_mf[i]) + w_mlp . h3[pair])
__global__ void __launch_bounds__(256) neumf_pair_head_kernel(const float *Umf, const float *Imf, int64_t ldt, int f, int u0, int n_ub,
                                                              int n_items, const float *h3, int64_t ldh, const float *wp,
                                                              const float *bp, float *prob, int64_t ldpr) {
    const int lane = threadIdx.x & 31;
    int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5, total = (int64_t)n_ub * n_items;
    for (; w < total; w += nw) {
        const int it = (int)(w % n_items), ub = (int)(w / n_items);
        float part = 0.f;
        for (int c = lane; c < 2 * f; c += 32)
            part += wp[c] * (c < f ? Umf[(int64_t)(u0 + ub) * ldt + c] * Imf[(int64_t)it * ldt + c] : h3[w * ldh + c - f]);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
        if (lane == 0) prob[(int64_t)ub * ldpr + it] = 1.f / (1.f + __expf(-(part + bp[0])));
    }
}

static inline unsigned ngrid(int64_t threads) {
    int64_t g = (threads + 255) / 256; const int64_t cap = (int64_t)sm_count() * 8;
    return (unsigned)(g > cap ? cap : (g < 1 ? 1 : g));
}

}  // namespace eb

using namespace eb;

extern "C" int eb_neumf_gather(const float *Umf, const float *Imf, const float *Umlp, const float *Imlp, int f, int64_t ldt,
                               const int32_t *u, const int32_t *it, int64_t n, float *x0, int64_t ldx, float *pm, int64_t ldp,
                               void *stream) {
    EB_ARG(Umf && Imf && Umlp && Imlp && u && it && x0 && pm && f >= 4 && f % 4 == 0 && f <= 128 && ldt % 4 == 0 && ldx % 4 == 0 && ldp % 4 == 0,
           "bad argument (f must be a multiple of 4, <= 128)");
    if (n <= 0) return EB_OK;
    neumf_gather_kernel<<<ngrid(n * 32), 256, 0, (cudaStream_t)stream>>>(Umf, Imf, Umlp, Imlp, f, ldt, u, it, n, x0, ldx, pm, ldp);
    EB_CUDA(cudaGetLastError());
    return EB_OK;
}

extern "C" int eb_neumf_head_norm(const float *pm, int64_t ldp, const float *h3, int64_t ldh, int f, const float *wp, const float *bp,
                                  const float *label, int64_t n, int64_t mean_over, float *dpm, float *dh3, float *dwp, float *dbp,
                                  double *loss, float *prob_out, void *stream) {
    EB_ARG(pm && h3 && wp && bp && f >= 1 && f <= 128 && n >= 0 && mean_over >= 1, "bad argument");
    EB_ARG(!label || (dpm && dh3 && dwp && dbp), "training mode needs the gradient outputs");
    if (n == 0) return EB_OK;
    neumf_head_kernel<<<ngrid(n * 32), 256, 0, (cudaStream_t)stream>>>(pm, ldp, h3, ldh, f, wp, bp, label, n, dpm, dh3, dwp, dbp, loss,
                                                                        prob_out, 1.f / (float)mean_over);
    EB_CUDA(cudaGetLastError());
    return EB_OK;
}

extern "C" int eb_neumf_head(const float *pm, int64_t ldp, const float *h3, int64_t ldh, int f, const float *wp, const float *bp,
                             const float *label, int64_t n, float *dpm, float *dh3, float *dwp, float *dbp, double *loss,
                             float *prob_out, void *stream) {
    return eb_neumf_head_norm(pm, ldp, h3, ldh, f, wp, bp, label, n, n > 0 ? n : 1, dpm, dh3, dwp, dbp, loss, prob_out, stream);
}

extern "C" int eb_relu_bwd_copy(const float *dout, const float *out, float *dpre, int64_t n, void *copy_bf16, void *stream) {
    EB_ARG(dout && out && dpre && n >= 0, "bad argument");
    if (n == 0) return EB_OK;
    relu_bwd_kernel<<<ngrid(n), 256, 0, (cudaStream_t)stream>>>(dout, out, dpre, n, (__nv_bfloat16 *)copy_bf16);
    EB_CUDA(cudaGetLastError());
    return EB_OK;
}

extern "C" int eb_relu_bwd(const float *dout, const float *out, float *dpre, int64_t n, void *stream) {
    return eb_relu_bwd_copy(dout, out, dpre, n, nullptr, stream);
}

extern "C" int eb_neumf_scatter(const float *Umf, const float *Imf, int f, int64_t ldt, const int32_t *u, const int32_t *it, int64_t n,
                                const float *dpm, int64_t ldp, const float *dx0, int64_t ldx, float *dUmf, float *dImf, float *dUmlp,
                                float *dImlp, void *stream) {
    EB_ARG(Umf && Imf && u && it && dpm && dx0 && dUmf && dImf && dUmlp && dImlp && f % 4 == 0, "bad argument");
    if (n <= 0) return EB_OK;
    neumf_scatter_kernel<<<ngrid(n * 32), 256, 0, (cudaStream_t)stream>>>(Umf, Imf, f, ldt, u, it, n, dpm, ldp, dx0, ldx, dUmf, dImf, dUmlp,
                                                                           dImlp);
    EB_CUDA(cudaGetLastError());
    return EB_OK;
}

extern "C" int eb_neumf_sample(int32_t n_users, int32_t n_items, const int64_t *indptr, const int32_t *indices, int m, uint64_t seed,
                               int64_t total, int32_t *out_u, int32_t *out_i, float *out_label, void *stream) {
    EB_ARG(indptr && indices && out_u && out_i && out_label && n_users >= 1 && n_items >= 2 && m >= 0 && total >= 0, "bad argument");
    if (total == 0) return EB_OK;
    neumf_sample_kernel<<<ngrid(total), 256, 0, (cudaStream_t)stream>>>(n_users, n_items, indptr, indices, m, seed, total, out_u, out_i,
                                                                         out_label);
    EB_CUDA(cudaGetLastError());
    return EB_OK;
}

extern "C" int eb_neumf_pair_h1(const float *Au, int64_t ldau, const float *Ai, int64_t ldai, const float *b1, int n_ub, int n_items,
                                int h1, void *out_bf16, int64_t ldo, void *stream) {
    EB_ARG(Au && Ai && b1 && out_bf16 && n_ub >= 1 && n_items >= 1 && h1 % 8 == 0 && ldo % 8 == 0 && ldau % 4 == 0 && ldai % 4 == 0, "bad argument");
    neumf_pair_h1_kernel<<<ngrid((int64_t)n_ub * n_items * (h1 / 4)), 256, 0, (cudaStream_t)stream>>>(Au, ldau, Ai, ldai, b1, n_ub, n_items,
                                                                                                       h1, (__nv_bfloat16 *)out_bf16, ldo);
    EB_CUDA(cudaGetLastError());
    return EB_OK;
}

extern "C" int eb_neumf_pair_h1_f32(const float *Au, int64_t ldau, const float *Ai, int64_t ldai, const float *b1, int n_ub, int n_items,
                                    int h1, float *out, int64_t ldo, void *stream) {
    EB_ARG(Au && Ai && b1 && out && n_ub >= 1 && n_items >= 1 && h1 >= 1 && ldo >= h1, "bad argument");
    neumf_pair_h1_f32_kernel<<<ngrid((int64_t)n_ub * n_items * h1), 256, 0, (cudaStream_t)stream>>>(Au, ldau, Ai, ldai, b1, n_ub, n_items, h1,
                                                                                                     out, ldo);
    EB_CUDA(cudaGetLastError());
    return EB_OK;
}

extern "C" int eb_neumf_pair_head(const float *Umf, const float *Imf, int64_t ldt, int f, int u0, int n_ub, int n_items, const float *h3,
                                  int64_t ldh, const float *wp, const float *bp, float *prob, int64_t ldpr, void *stream) {
    EB_ARG(Umf && Imf && h3 && wp && bp && prob && n_ub >= 1 && n_items >= 1, "bad argument");
    neumf_pair_head_kernel<<<ngrid((int64_t)n_ub * n_items * 32), 256, 0, (cudaStream_t)stream>>>(Umf, Imf, ldt, f, u0, n_ub, n_items, h3, ldh,
                                                                                                   wp, bp, prob, ldpr);
    EB_CUDA(cudaGetLastError());
    return EB_OK;
}
