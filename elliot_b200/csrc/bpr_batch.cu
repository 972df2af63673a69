// bpr_batch.cu — mini-batch BPR-MF with Adam (the reference's TensorFlow variant).
//
// Replaces BPRMF_batch_model.call/train_step
// (elliot/recommender/latent_factor_models/BPRMF_batch/BPRMF_batch_model.py:46-80):
//   x = Bi[i] + <Gu[u], Gi[i]>;  diff = clip(x_pos - x_neg, -80, 1e8)
//   loss = sum softplus(-diff) + l_w (|Gu[u]|^2 + |Gi[pos]|^2 + |Gi[neg]|^2)/2
//          + l_b Bi[pos]^2/2 + l_b Bi[neg]^2/2/10          (tf.nn.l2_loss = sum x^2 / 2, per batch ROW)
//   grads are IndexedSlices (duplicates summed); optimizer = Keras Adam, whose sparse path in
//   TF 2.3 decays m and v and moves the variable for ALL rows every step
//   (OptimizerV2 Adam._resource_apply_sparse: m*=b1; v*=b2; scatter_add; var -= lr_t m/(sqrt(v)+eps)).
// Two kernels: (1) gather-score-grad with 128-bit vector atomics into dense gradient tables,
// (2) one streaming dense Adam pass per table that also clears the gradient.
// tensorflow==2.3.2 cannot be executed in the build container: parity with TF is UNPINNED; the
// checker is the numpy restatement oracle/tf_models.py::bprmf_batch_step.
#include <cuda_bf16.h>

#include "common.cuh"

namespace eb {

struct BatchGradParams {
    const float *Gu, *Gi, *Bi;
    float *dGu, *dGi, *dBi;
    int ld;
    const int32_t *tu, *ti, *tj;
    int64_t n;
    float l_w, l_b;
    double *loss;
};

__device__ __forceinline__ void red4(float *p, float4 v) {
    asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

template <int DP>
__global__ void __launch_bounds__(256, 4) bpr_batch_grad_kernel(const BatchGradParams p) {
    constexpr int NV = DP / 4, G = NV >= 32 ? 32 : NV, VPL = NV / G;
    const int lane = threadIdx.x & 31, gl = lane % G, gbase = lane - gl;
    const int64_t warp_id = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t ld = p.ld;
    float loss_acc = 0.f;
    for (int64_t tile = warp_id; tile * 32 < p.n; tile += nwarps) {
        const int64_t t = tile * 32 + lane;
        int u = -1, i = 0, j = 0;
        if (t < p.n) { u = __ldg(p.tu + t); i = __ldg(p.ti + t); j = __ldg(p.tj + t); }
#pragma unroll 1
        for (int s = 0; s < G; s++) {
            const int cu = __shfl_sync(0xffffffffu, u, gbase + s), ci = __shfl_sync(0xffffffffu, i, gbase + s),
                      cj = __shfl_sync(0xffffffffu, j, gbase + s);
            float4 a[VPL], vi[VPL], vj[VPL];
            float part_i = 0.f, part_j = 0.f, sq = 0.f;
            if (cu >= 0) {
#pragma unroll
                for (int v = 0; v < VPL; v++) {
                    a[v] = reinterpret_cast<const float4 *>(p.Gu + (int64_t)cu * ld)[v * G + gl];
                    vi[v] = reinterpret_cast<const float4 *>(p.Gi + (int64_t)ci * ld)[v * G + gl];
                    vj[v] = reinterpret_cast<const float4 *>(p.Gi + (int64_t)cj * ld)[v * G + gl];
                    part_i += a[v].x * vi[v].x + a[v].y * vi[v].y + a[v].z * vi[v].z + a[v].w * vi[v].w;
                    part_j += a[v].x * vj[v].x + a[v].y * vj[v].y + a[v].z * vj[v].z + a[v].w * vj[v].w;
                    sq += a[v].x * a[v].x + a[v].y * a[v].y + a[v].z * a[v].z + a[v].w * a[v].w +
                          vi[v].x * vi[v].x + vi[v].y * vi[v].y + vi[v].z * vi[v].z + vi[v].w * vi[v].w +
                          vj[v].x * vj[v].x + vj[v].y * vj[v].y + vj[v].z * vj[v].z + vj[v].w * vj[v].w;
                }
            }
#pragma unroll
            for (int off = G / 2; off > 0; off >>= 1) {
                part_i += __shfl_xor_sync(0xffffffffu, part_i, off);
                part_j += __shfl_xor_sync(0xffffffffu, part_j, off);
                sq += __shfl_xor_sync(0xffffffffu, sq, off);
            }
            if (cu >= 0) {
                const float bi = p.Bi[ci], bj = p.Bi[cj];
                const float raw = (bi + part_i) - (bj + part_j);
                const float diff = fminf(fmaxf(raw, -80.f), 1e8f);            // BPRMF_batch_model.py:64
                const bool open = raw > -80.f && raw < 1e8f;                   // clip_by_value passes gradient only inside
                const float sg = 1.f / (1.f + __expf(diff));                  // sigmoid(-diff)
                const float g = open ? -sg : 0.f;                              // dL/d(x_pos - x_neg)
                if (gl == 0)
                    loss_acc += fmaxf(-diff, 0.f) + __logf(1.f + __expf(-fabsf(diff))) + 0.5f * p.l_w * sq +
                                0.5f * p.l_b * bi * bi + 0.05f * p.l_b * bj * bj;
#pragma unroll
                for (int v = 0; v < VPL; v++) {
                    const int e = (v * G + gl) * 4;
                    float4 du, di, dj;
                    du.x = g * (vi[v].x - vj[v].x) + p.l_w * a[v].x; du.y = g * (vi[v].y - vj[v].y) + p.l_w * a[v].y;
                    du.z = g * (vi[v].z - vj[v].z) + p.l_w * a[v].z; du.w = g * (vi[v].w - vj[v].w) + p.l_w * a[v].w;
                    di.x = g * a[v].x + p.l_w * vi[v].x; di.y = g * a[v].y + p.l_w * vi[v].y;
                    di.z = g * a[v].z + p.l_w * vi[v].z; di.w = g * a[v].w + p.l_w * vi[v].w;
                    dj.x = -g * a[v].x + p.l_w * vj[v].x; dj.y = -g * a[v].y + p.l_w * vj[v].y;
                    dj.z = -g * a[v].z + p.l_w * vj[v].z; dj.w = -g * a[v].w + p.l_w * vj[v].w;
                    red4(p.dGu + (int64_t)cu * ld + e, du);
                    red4(p.dGi + (int64_t)ci * ld + e, di);
                    red4(p.dGi + (int64_t)cj * ld + e, dj);
                }
                if (gl == 0) {
                    atomicAdd(p.dBi + ci, g + p.l_b * bi);
                    atomicAdd(p.dBi + cj, -g + 0.1f * p.l_b * bj);
                }
            }
        }
    }
    if (p.loss) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) loss_acc += __shfl_xor_sync(0xffffffffu, loss_acc, off);
        if (lane == 0 && loss_acc != 0.f) atomicAdd(p.loss, (double)loss_acc);
    }
}

// Keras Adam, dense form (what TF 2.3 does to every row of a variable even for sparse grads)
__global__ void __launch_bounds__(256) adam_dense_kernel(float *__restrict__ var, float *__restrict__ m,
                                                         float *__restrict__ v, float *__restrict__ grad, int64_t n4,
                                                         float b1, float b2, float lr_t, float eps,
                                                         __nv_bfloat162 *__restrict__ copy_bf16) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n4; i += stride) {
        float4 g = reinterpret_cast<float4 *>(grad)[i], mm = reinterpret_cast<float4 *>(m)[i],
               vv = reinterpret_cast<float4 *>(v)[i], x = reinterpret_cast<float4 *>(var)[i];
        mm.x = b1 * mm.x + (1.f - b1) * g.x; mm.y = b1 * mm.y + (1.f - b1) * g.y;
        mm.z = b1 * mm.z + (1.f - b1) * g.z; mm.w = b1 * mm.w + (1.f - b1) * g.w;
        vv.x = b2 * vv.x + (1.f - b2) * g.x * g.x; vv.y = b2 * vv.y + (1.f - b2) * g.y * g.y;
        vv.z = b2 * vv.z + (1.f - b2) * g.z * g.z; vv.w = b2 * vv.w + (1.f - b2) * g.w * g.w;
        x.x -= lr_t * mm.x / (sqrtf(vv.x) + eps); x.y -= lr_t * mm.y / (sqrtf(vv.y) + eps);
        x.z -= lr_t * mm.z / (sqrtf(vv.z) + eps); x.w -= lr_t * mm.w / (sqrtf(vv.w) + eps);
        reinterpret_cast<float4 *>(m)[i] = mm; reinterpret_cast<float4 *>(v)[i] = vv;
        reinterpret_cast<float4 *>(var)[i] = x;
        if (copy_bf16) {                               // the tensor-core operand copy of the new weights, same layout
            copy_bf16[2 * i] = __floats2bfloat162_rn(x.x, x.y);
            copy_bf16[2 * i + 1] = __floats2bfloat162_rn(x.z, x.w);
        }
        reinterpret_cast<float4 *>(grad)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

}  // namespace eb

using namespace eb;

extern "C" int eb_bpr_batch_grad_f32(const float *Gu, const float *Gi, const float *Bi, float *dGu, float *dGi, float *dBi,
                                     int d, int ld, const int32_t *tu, const int32_t *ti, const int32_t *tj, int64_t n,
                                     float l_w, float l_b, double *loss, void *stream) {
    EB_ARG(Gu && Gi && Bi && dGu && dGi && dBi, "null table pointer");
    EB_ARG(d >= 1 && ld >= d, "need 1 <= d <= ld");
    if (n <= 0) return EB_OK;
    EB_ARG(tu && ti && tj, "null triple arrays");
    BatchGradParams p{Gu, Gi, Bi, dGu, dGi, dBi, ld, tu, ti, tj, n, l_w, l_b, loss};
    int64_t grid = ((n + 31) / 32 + 7) / 8;
    const int64_t cap = (int64_t)sm_count() * 4;
    if (grid > cap) grid = cap;
    cudaStream_t st = (cudaStream_t)stream;
    switch (ld) {
        case 8: bpr_batch_grad_kernel<8><<<(unsigned)grid, 256, 0, st>>>(p); break;
        case 16: bpr_batch_grad_kernel<16><<<(unsigned)grid, 256, 0, st>>>(p); break;
        case 32: bpr_batch_grad_kernel<32><<<(unsigned)grid, 256, 0, st>>>(p); break;
        case 64: bpr_batch_grad_kernel<64><<<(unsigned)grid, 256, 0, st>>>(p); break;
        case 128: bpr_batch_grad_kernel<128><<<(unsigned)grid, 256, 0, st>>>(p); break;
        case 256: bpr_batch_grad_kernel<256><<<(unsigned)grid, 256, 0, st>>>(p); break;
        default: return set_err(EB_ERR_ARG, "row stride ld=%d must be one of 8,16,32,64,128,256 floats", ld);
    }
    EB_CUDA(cudaGetLastError());
    return EB_OK;
}

extern "C" int eb_adam_dense_f32(float *var, float *m, float *v, float *grad, int64_t n, float lr, float beta1, float beta2,
                                 float eps, int64_t step, void *stream) {
    return eb_adam_dense_copy_f32(var, m, v, grad, n, lr, beta1, beta2, eps, step, nullptr, stream);
}

extern "C" int eb_adam_dense_copy_f32(float *var, float *m, float *v, float *grad, int64_t n, float lr, float beta1, float beta2,
                                      float eps, int64_t step, void *copy_bf16, void *stream) {
    EB_ARG(var && m && v && grad && n >= 0 && (n % 4) == 0 && step >= 1, "bad argument (n must be a multiple of 4, step >= 1)");
    EB_ARG(!copy_bf16 || ((uintptr_t)copy_bf16 % 4) == 0, "bf16 copy must be 4-byte aligned");
    if (n == 0) return EB_OK;
    const double lr_t = (double)lr * sqrt(1.0 - pow((double)beta2, (double)step)) / (1.0 - pow((double)beta1, (double)step));
    int64_t grid = (n / 4 + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (grid > cap) grid = cap;
    adam_dense_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(var, m, v, grad, n / 4, beta1, beta2, (float)lr_t, eps,
                                                                        (__nv_bfloat162 *)copy_bf16);
    EB_CUDA(cudaGetLastError());
    return EB_OK;
}

// ---- replicated-table reconciliation (multi-GPU): delta = cur - prev ; cur = prev = prev + sum(delta)
namespace eb {
__global__ void __launch_bounds__(256) delta_kernel(const float4 *__restrict__ cur, const float4 *__restrict__ prev,
                                                    float4 *__restrict__ out, int64_t n4) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 a = cur[i], b = prev[i];
        out[i] = make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w);
    }
}
__global__ void __launch_bounds__(256) apply_delta_kernel(float4 *__restrict__ cur, float4 *__restrict__ prev,
                                                          const float4 *__restrict__ sum, int64_t n4, float scale) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 b = prev[i], s = sum[i];
        const float4 r = make_float4(b.x + scale * s.x, b.y + scale * s.y, b.z + scale * s.z, b.w + scale * s.w);
        cur[i] = r; prev[i] = r;
    }
}
// overlapped variant: cur already contains newer local updates; add what the OTHER ranks contributed
//   cur += (scale*sum - local);  prev += scale*sum
// cur is updated with vector atomics: a training kernel on another stream may be adding to the same rows.
__global__ void __launch_bounds__(256) apply_delta_late_kernel(float *__restrict__ cur, float4 *__restrict__ prev,
                                                               const float4 *__restrict__ sum, const float4 *__restrict__ local,
                                                               int64_t n4, float scale) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 s = sum[i];
        const float4 l = local[i];
        s.x *= scale; s.y *= scale; s.z *= scale; s.w *= scale;
        float4 b = prev[i];
        red_add_v4(cur + 4 * i, make_float4(s.x - l.x, s.y - l.y, s.z - l.z, s.w - l.w));
        b.x += s.x; b.y += s.y; b.z += s.z; b.w += s.w;
        prev[i] = b;
    }
}
}  // namespace eb

extern "C" int eb_table_apply_delta_late_f32(float *cur, float *prev, const float *delta_sum, const float *delta_local, int64_t n,
                                             float scale, void *stream) {
    EB_ARG(cur && prev && delta_sum && delta_local && n >= 0 && n % 4 == 0, "bad argument (n must be a multiple of 4)");
    if (n == 0) return EB_OK;
    int64_t grid = (n / 4 + 255) / 256; const int64_t cap = (int64_t)eb::sm_count() * 8; if (grid > cap) grid = cap;
    eb::apply_delta_late_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(cur, (float4 *)prev, (const float4 *)delta_sum,
                                                                                  (const float4 *)delta_local, n / 4, scale);
    EB_CUDA(cudaGetLastError());
    return EB_OK;
}

extern "C" int eb_table_delta_f32(const float *cur, const float *prev, float *delta, int64_t n, void *stream) {
    EB_ARG(cur && prev && delta && n >= 0 && n % 4 == 0, "bad argument (n must be a multiple of 4)");
    if (n == 0) return EB_OK;
    int64_t grid = (n / 4 + 255) / 256; const int64_t cap = (int64_t)eb::sm_count() * 8; if (grid > cap) grid = cap;
    eb::delta_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>((const float4 *)cur, (const float4 *)prev, (float4 *)delta, n / 4);
    EB_CUDA(cudaGetLastError());
    return EB_OK;
}

extern "C" int eb_table_apply_delta_f32(float *cur, float *prev, const float *delta_sum, int64_t n, float scale, void *stream) {
    EB_ARG(cur && prev && delta_sum && n >= 0 && n % 4 == 0, "bad argument (n must be a multiple of 4)");
    if (n == 0) return EB_OK;
    int64_t grid = (n / 4 + 255) / 256; const int64_t cap = (int64_t)eb::sm_count() * 8; if (grid > cap) grid = cap;
    eb::apply_delta_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>((float4 *)cur, (float4 *)prev, (const float4 *)delta_sum, n / 4, scale);
    EB_CUDA(cudaGetLastError());
    return EB_OK;
}
