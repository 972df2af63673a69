// mf2020.cu — pointwise logistic matrix factorisation ("NCF vs MF revisited"), the sibling of the BPR step that
// shares its gather / dot / scatter shape.
//
// Replaces MFModel.train_step (elliot/recommender/latent_factor_models/MF2020/MF_model.py:80-112):
//   pred = gb + ub[u] + ib[i] + U[u].V[i];  sigmoid/loss split at pred > 0 (:94-101);  grad = rating - sigmoid
//   U[u] += lr (grad V[i] - reg U[u]);  V[i] += lr (grad U'[u] - reg V[i])   (U' = the row just written: uf_ is a view)
//   ub[u] += lr (grad - reg ub);  ib[i] += lr (grad - reg ib);  gb += lr (grad - reg gb)
//
// Exact mode (fp64).  Every sample reads and writes the GLOBAL bias, so the reference's order is a single
// dependency chain — there is no row-level parallelism to recover as in BPR.  One warp walks the sample list in
// order; lane l owns elements l, l+32, ... of both rows (so every address is always read and written by the same
// thread and plain program order gives sequential consistency, no fences or counters).  Latency is hidden by
// software pipelining: the rows of sample t+1 are loaded while sample t computes and are patched from registers
// when t+1 touches a row t has just rewritten; rows further ahead are pulled into L2 with prefetch hints.
//
// Throughput mode (fp32, Hogwild): one launch over every positive of the train CSR plus m uniform negatives each
// (custom_sampler_rendle.py:29-85: negatives are NOT rejected against the train set), visited in a
// pseudo-random order, 128-bit row loads, vector atomics; the global bias moves once per warp.
#include "common.cuh"

namespace eb {

struct MfSeqParams {
    double *U, *V, *ub, *ib, *gb;
    int d, ld;
    const int32_t *su, *si, *sr;
    int64_t n, batch;
    double lr, reg;
    double *batch_loss;
};


template <int NE>
__global__ void __launch_bounds__(32) mf_pointwise_seq_kernel(const MfSeqParams p) {
    const int lane = threadIdx.x;
    if (p.n <= 0) return;
    double gb = *p.gb;
    int u = __ldg(p.su), i = __ldg(p.si);
    double r = (double)__ldg(p.sr);
    double a[NE], v[NE];
    {
        const double *pu = p.U + (int64_t)u * p.ld, *pv = p.V + (int64_t)i * p.ld;
#pragma unroll
        for (int e = 0; e < NE; e++) {
            const int k = lane + 32 * e;
            a[e] = k < p.d ? __ldcg(pu + k) : 0.0;
            v[e] = k < p.d ? __ldcg(pv + k) : 0.0;
        }
    }
    double bu = __ldcg(p.ub + u), bi = __ldcg(p.ib + i);       // every lane keeps (and later stores) its own copy
    double lacc = 0.0;
    constexpr int PF = 8;
    for (int64_t t = 0; t < p.n; ++t) {
        // ---- stage sample t+1 (its loads fly while sample t computes)
        const bool has = t + 1 < p.n;
        int un = u, in = i;
        double rn = 0.0, bun = 0.0, bin = 0.0, an[NE], vn[NE];
        if (has) {
            un = __ldg(p.su + t + 1); in = __ldg(p.si + t + 1); rn = (double)__ldg(p.sr + t + 1);
            const double *pu = p.U + (int64_t)un * p.ld, *pv = p.V + (int64_t)in * p.ld;
#pragma unroll
            for (int e = 0; e < NE; e++) {
                const int k = lane + 32 * e;
                an[e] = k < p.d ? __ldcg(pu + k) : 0.0;
                vn[e] = k < p.d ? __ldcg(pv + k) : 0.0;
            }
            bun = __ldcg(p.ub + un); bin = __ldcg(p.ib + in);
        }
        if (t + PF < p.n && lane * 16 < p.d) {                  // one hint per 128-byte line of the two rows
            prefetch_l2(p.U + (int64_t)__ldg(p.su + t + PF) * p.ld + lane * 16);
            prefetch_l2(p.V + (int64_t)__ldg(p.si + t + PF) * p.ld + lane * 16);
        }
        // ---- sample t (MF_model.py:84-111)
        double dot = 0.0;
#pragma unroll
        for (int e = 0; e < NE; e++) dot = __dadd_rn(dot, __dmul_rn(a[e], v[e]));
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) dot = __dadd_rn(dot, __shfl_xor_sync(0xffffffffu, dot, off));
        const double pred = __dadd_rn(__dadd_rn(__dadd_rn(gb, bu), bi), dot);
        double sig, loss;
        if (pred > 0) {                                         // warp-uniform: every lane holds the same pred
            const double ope = 1.0 + exp(-pred);
            sig = 1.0 / ope;
            loss = log(ope) + (1.0 - r) * pred;
        } else {
            const double ex = exp(pred);
            sig = ex / (1.0 + ex);
            loss = -r * pred + log(1.0 + ex);
        }
        const double grad = r - sig;
        double *pu = p.U + (int64_t)u * p.ld, *pv = p.V + (int64_t)i * p.ld;
#pragma unroll
        for (int e = 0; e < NE; e++) {
            const int k = lane + 32 * e;
            const double un_ = __dadd_rn(a[e], __dmul_rn(p.lr, __dadd_rn(__dmul_rn(grad, v[e]), -__dmul_rn(p.reg, a[e]))));
            const double vn_ = __dadd_rn(v[e], __dmul_rn(p.lr, __dadd_rn(__dmul_rn(grad, un_), -__dmul_rn(p.reg, v[e]))));
            a[e] = un_; v[e] = vn_;
            if (k < p.d) { __stcg(pu + k, un_); __stcg(pv + k, vn_); }
        }
        bu = __dadd_rn(bu, __dmul_rn(p.lr, __dadd_rn(grad, -__dmul_rn(p.reg, bu))));
        bi = __dadd_rn(bi, __dmul_rn(p.lr, __dadd_rn(grad, -__dmul_rn(p.reg, bi))));
        gb = __dadd_rn(gb, __dmul_rn(p.lr, __dadd_rn(grad, -__dmul_rn(p.reg, gb))));
        __stcg(p.ub + u, bu); __stcg(p.ib + i, bi);             // all lanes, same value: each lane re-reads its own store
        lacc += loss;
        if (p.batch_loss && ((t + 1) % p.batch == 0 || !has)) {
            if (lane == 0) p.batch_loss[t / p.batch] = lacc;
            lacc = 0.0;
        }
        // ---- hand over: rows that sample t rewrote replace the copies staged before the write
        if (has) {
            if (un == u) {
#pragma unroll
                for (int e = 0; e < NE; e++) an[e] = a[e];
                bun = bu;
            }
            if (in == i) {
#pragma unroll
                for (int e = 0; e < NE; e++) vn[e] = v[e];
                bin = bi;
            }
#pragma unroll
            for (int e = 0; e < NE; e++) { a[e] = an[e]; v[e] = vn[e]; }
            u = un; i = in; r = rn; bu = bun; bi = bin;
        }
    }
    if (lane == 0) *p.gb = gb;
}

// ---------------------------------------------------------------- throughput mode (fp32, Hogwild)
struct MfHogParams {
    float *U, *V, *ub, *ib, *gb;
    int ld;
    const int32_t *pos_u, *pos_i;
    int64_t n_pos, n;          // n = n_pos * (1 + m) samples per epoch
    int64_t s_begin, s_end;    // this launch covers epoch positions [s_begin, s_end)
    double *gb_work;           // {sum grad, sum sig(1-sig), samples} of the launch
    int m;
    int32_t n_items;
    uint64_t seed, first, mul, add;
    float lr, reg;
    double *loss;
    int32_t *out_u, *out_i, *out_r;
};

// sample s of the epoch: a fixed affine permutation of [0, n) stands in for random.sample (custom_sampler_rendle.py:80),
// slot 0 of each positive is the positive itself, slots 1..m are uniform items (no rejection, :66-69)
__device__ __forceinline__ void mf_sample(const MfHogParams &p, int64_t s, int &u, int &i, float &r) {
    const uint64_t sp = ((uint64_t)s * p.mul + p.add) % (uint64_t)p.n;      // mul < 2^23, n < 2^40: no overflow
    const int64_t pos = (int64_t)(sp / (uint64_t)(1 + p.m));
    const int q = (int)(sp % (uint64_t)(1 + p.m));
    u = __ldg(p.pos_u + pos);
    if (q == 0) { i = __ldg(p.pos_i + pos); r = 1.f; return; }
    uint32_t x[4];
    Philox::gen(p.seed, p.first + sp, 0u, x);
    i = (int)bounded(x[0], (uint32_t)p.n_items);
    r = 0.f;
}

template <int DP>
__global__ void __launch_bounds__(256) mf_hogwild_kernel(const MfHogParams p) {
    constexpr int NV = DP / 4;
    constexpr int G = NV >= 32 ? 32 : NV;
    constexpr int VPL = NV / G;
    constexpr int UNR = G >= 4 ? 4 : G;
    const int lane = threadIdx.x & 31;
    const int gl = lane % G, gbase = lane - gl;
    const int64_t warp_id = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t ld = p.ld;
    const float gb0 = *p.gb;                                      // stale within the launch (Hogwild)
    float loss_acc = 0.f, gsum = 0.f, hsum = 0.f, gcnt = 0.f;
    for (int64_t tile = warp_id; p.s_begin + tile * 32 < p.s_end; tile += nwarps) {
        const int64_t s = p.s_begin + tile * 32 + lane;
        int u = -1, i = 0;
        float r = 0.f;
        if (s < p.s_end) {
            mf_sample(p, s, u, i, r);
            if (p.out_u) { p.out_u[s] = u; p.out_i[s] = i; p.out_r[s] = (int32_t)r; }
        }
#pragma unroll 1
        for (int s0 = 0; s0 < G; s0 += UNR) {
            float4 ru[UNR][VPL], rv[UNR][VPL];
            float bu[UNR], bi[UNR], rr[UNR];
            int cu[UNR], ci[UNR];
#pragma unroll
            for (int q = 0; q < UNR; q++) {
                cu[q] = __shfl_sync(0xffffffffu, u, gbase + s0 + q);
                ci[q] = __shfl_sync(0xffffffffu, i, gbase + s0 + q);
                rr[q] = __shfl_sync(0xffffffffu, r, gbase + s0 + q);
                if (cu[q] >= 0) {
                    const float4 *pu = reinterpret_cast<const float4 *>(p.U + (int64_t)cu[q] * ld);
                    const float4 *pv = reinterpret_cast<const float4 *>(p.V + (int64_t)ci[q] * ld);
#pragma unroll
                    for (int v = 0; v < VPL; v++) { ru[q][v] = pu[v * G + gl]; rv[q][v] = pv[v * G + gl]; }
                    bu[q] = p.ub[cu[q]]; bi[q] = p.ib[ci[q]];
                }
            }
#pragma unroll
            for (int q = 0; q < UNR; q++) {
                float part = 0.f;
                if (cu[q] >= 0) {
#pragma unroll
                    for (int v = 0; v < VPL; v++)
                        part += ru[q][v].x * rv[q][v].x + ru[q][v].y * rv[q][v].y + ru[q][v].z * rv[q][v].z + ru[q][v].w * rv[q][v].w;
                }
#pragma unroll
                for (int off = G / 2; off > 0; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
                if (cu[q] >= 0) {
                    const float pred = gb0 + bu[q] + bi[q] + part;
                    const float e = __expf(-fabsf(pred));
                    const float sig = pred > 0.f ? __fdividef(1.f, 1.f + e) : __fdividef(e, 1.f + e);   // MF_model.py:94-101
                    const float grad = rr[q] - sig;
                    if (gl == 0) {
                        loss_acc += fmaxf(pred, 0.f) - rr[q] * pred + __logf(1.f + e);
                        gsum += grad; hsum += sig * (1.f - sig); gcnt += 1.f;
                        red_add_f32(p.ub + cu[q], p.lr * (grad - p.reg * bu[q]));
                        red_add_f32(p.ib + ci[q], p.lr * (grad - p.reg * bi[q]));
                    }
                    float *pu = p.U + (int64_t)cu[q] * ld, *pv = p.V + (int64_t)ci[q] * ld;
#pragma unroll
                    for (int v = 0; v < VPL; v++) {
                        const float4 a = ru[q][v], b = rv[q][v];
                        float4 du, dv;
                        du.x = p.lr * (grad * b.x - p.reg * a.x); du.y = p.lr * (grad * b.y - p.reg * a.y);
                        du.z = p.lr * (grad * b.z - p.reg * a.z); du.w = p.lr * (grad * b.w - p.reg * a.w);
                        // the item row sees the UPDATED user row (view aliasing, MF_model.py:105-106)
                        dv.x = p.lr * (grad * (a.x + du.x) - p.reg * b.x); dv.y = p.lr * (grad * (a.y + du.y) - p.reg * b.y);
                        dv.z = p.lr * (grad * (a.z + du.z) - p.reg * b.z); dv.w = p.lr * (grad * (a.w + du.w) - p.reg * b.w);
                        const int el = (v * G + gl) * 4;
                        red_add_v4(pu + el, du);
                        red_add_v4(pv + el, dv);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        loss_acc += __shfl_xor_sync(0xffffffffu, loss_acc, off);
        gsum += __shfl_xor_sync(0xffffffffu, gsum, off);
        hsum += __shfl_xor_sync(0xffffffffu, hsum, off);
        gcnt += __shfl_xor_sync(0xffffffffu, gcnt, off);
    }
    if (lane == 0 && gcnt != 0.f) {
        atomicAdd(p.gb_work + 0, (double)gsum); atomicAdd(p.gb_work + 1, (double)hsum); atomicAdd(p.gb_work + 2, (double)gcnt);
        if (p.loss) atomicAdd(p.loss, (double)loss_acc);
    }
}

// The global bias is one scalar hit by EVERY sample: applying c stale gradients at once (gb += lr*sum) overshoots as soon
// as lr*c*sig' > 2.  The reference takes c tiny sequential steps gb += lr (g_s(gb) - reg gb); linearising the
// gradient around the launch's stale value, g_s(gb) ~ g_s(gb0) - sig'_s (gb - gb0), those steps integrate to
//   gb = gb0 + (gbar - reg gb0) / (hbar + reg) * (1 - exp(-lr (hbar + reg) c)),   gbar = sum g / c, hbar = sum sig' / c,
// which is the plain step for small lr*c and saturates at the launch's fixed point for large ones.
__global__ void mf_global_bias_finish_kernel(float *gb, double *work, float lr, float reg) {
    const double c = work[2];
    if (c > 0) {
        const double g0 = (double)*gb, gbar = work[0] / c, k = work[1] / c + (double)reg;
        const double x = (double)lr * k * c;
        const double gain = k > 1e-12 ? -expm1(-x) / k : (double)lr * c;
        *gb = (float)(g0 + (gbar - (double)reg * g0) * gain);
    }
    work[0] = 0; work[1] = 0; work[2] = 0;
}

template <int DP>
static int launch_mf_hogwild(const MfHogParams &p, cudaStream_t st) {
    int per_sm = 0;
    EB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mf_hogwild_kernel<DP>, 256, 0));
    if (per_sm < 1) per_sm = 1;
    const int64_t tiles = (p.s_end - p.s_begin + 31) / 32, want = (tiles + 7) / 8;
    int64_t grid = (int64_t)sm_count() * per_sm;
    if (want < grid) grid = want;
    if (grid < 1) grid = 1;
    mf_hogwild_kernel<DP><<<(unsigned)grid, 256, 0, st>>>(p);
    EB_CUDA(cudaGetLastError());
    mf_global_bias_finish_kernel<<<1, 1, 0, st>>>(p.gb, p.gb_work, p.lr, p.reg);
    EB_CUDA(cudaGetLastError());
    return EB_OK;
}

static uint64_t gcd_u64(uint64_t a, uint64_t b) { while (b) { uint64_t t = a % b; a = b; b = t; } return a; }

}  // namespace eb

extern "C" int eb_mf_pointwise_exact_f64(double *U, double *V, double *user_bias, double *item_bias, double *global_bias,
                                         int d, int ld, const int32_t *su, const int32_t *si, const int32_t *sr, int64_t n,
                                         double lr, double reg, int64_t batch, double *batch_loss, void *stream) {
    using namespace eb;
    EB_ARG(U && V && user_bias && item_bias && global_bias, "null table pointer");
    EB_ARG(d >= 1 && ld >= d && d <= 256, "need 1 <= d <= min(ld, 256) (d=%d ld=%d)", d, ld);
    EB_ARG(n >= 0 && batch >= 1, "bad n / batch");
    if (n == 0) return EB_OK;
    EB_ARG(su && si && sr, "null sample pointer");
    MfSeqParams p{U, V, user_bias, item_bias, global_bias, d, ld, su, si, sr, n, batch, lr, reg, batch_loss};
    cudaStream_t st = (cudaStream_t)stream;
    if (d <= 32) mf_pointwise_seq_kernel<1><<<1, 32, 0, st>>>(p);
    else if (d <= 64) mf_pointwise_seq_kernel<2><<<1, 32, 0, st>>>(p);
    else if (d <= 128) mf_pointwise_seq_kernel<4><<<1, 32, 0, st>>>(p);
    else mf_pointwise_seq_kernel<8><<<1, 32, 0, st>>>(p);
    EB_CUDA(cudaGetLastError());
    return EB_OK;
}

extern "C" int eb_mf_pointwise_step_f32(float *U, float *V, float *user_bias, float *item_bias, float *global_bias, int d, int ld,
                                        const int32_t *pos_u, const int32_t *pos_i, int64_t n_pos, int m, int32_t n_items,
                                        uint64_t seed, uint64_t epoch, int64_t first, int64_t count, float lr, float reg,
                                        double *loss, double *gb_work, int32_t *out_u, int32_t *out_i, int32_t *out_r,
                                        void *stream) {
    using namespace eb;
    EB_ARG(U && V && user_bias && item_bias && global_bias, "null table pointer");
    EB_ARG(d >= 1 && ld >= d, "need 1 <= d <= ld (d=%d ld=%d)", d, ld);
    EB_ARG(((uintptr_t)U % 16) == 0 && ((uintptr_t)V % 16) == 0, "tables must be 16-byte aligned");
    EB_ARG(n_pos >= 0 && m >= 0 && n_items >= 1, "bad n_pos / m / n_items");
    EB_ARG((out_u == nullptr) == (out_i == nullptr) && (out_u == nullptr) == (out_r == nullptr), "out_u/out_i/out_r go together");
    EB_ARG(gb_work, "gb_work (device double[4], zero-initialised) is required");
    const int64_t n_epoch = n_pos * (int64_t)(1 + m);
    EB_ARG(first >= 0 && count >= 0 && first + count <= n_epoch, "samples [%lld, %lld) outside the epoch of %lld",
           (long long)first, (long long)(first + count), (long long)n_epoch);
    if (count == 0) return EB_OK;
    EB_ARG(pos_u && pos_i, "null positives");
    MfHogParams p{};
    p.s_begin = first; p.s_end = first + count; p.gb_work = gb_work;
    p.U = U; p.V = V; p.ub = user_bias; p.ib = item_bias; p.gb = global_bias; p.ld = ld;
    p.pos_u = pos_u; p.pos_i = pos_i; p.n_pos = n_pos; p.m = m; p.n = n_pos * (int64_t)(1 + m); p.n_items = n_items;
    p.seed = seed; p.first = epoch * (uint64_t)p.n;
    // affine visiting order s -> (s*mul + add) mod n, mul coprime with n, re-drawn every epoch
    EB_ARG(p.n < (1ll << 40), "more than 2^40 samples per epoch");
    uint64_t mul = (((0x9E3779B97F4A7C15ull ^ (seed * 0xD1B54A32D192ED03ull + epoch)) >> 20) & 0x3fffffull) | 0x400001ull;   // odd, in [2^22, 2^23)
    while (gcd_u64(mul, (uint64_t)p.n) != 1) mul += 2;
    p.mul = mul; p.add = (seed * 0x2545F4914F6CDD1Dull + epoch * 0x632BE59BD9B4E019ull) % (uint64_t)p.n;
    p.lr = lr; p.reg = reg; p.loss = loss; p.out_u = out_u; p.out_i = out_i; p.out_r = out_r;
    cudaStream_t st = (cudaStream_t)stream;
    switch (ld) {
        case 8: return launch_mf_hogwild<8>(p, st);
        case 16: return launch_mf_hogwild<16>(p, st);
        case 32: return launch_mf_hogwild<32>(p, st);
        case 64: return launch_mf_hogwild<64>(p, st);
        case 128: return launch_mf_hogwild<128>(p, st);
        case 256: return launch_mf_hogwild<256>(p, st);
        default: return set_err(EB_ERR_ARG, "row stride ld=%d must be one of 8,16,32,64,128,256 floats", ld);
    }
}
