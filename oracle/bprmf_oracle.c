/*
 * oracle/bprmf_oracle.c — CPU restatement of the reference's BPRMF hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under elliot_b200/ may import, link or
 * execute this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference leg do, and there only as the checker.
 *
 * Every function cites the reference line range it restates (paths relative to
 * the upstream tree, sisinflab/elliot v0.3.1).  The restatement is pinned
 * against the reference itself executed in the build container
 * (oracle/gen_golden.py imports custom_sampler.py / BPRMF_model.py by file
 * path and writes tests/golden/*.npz) — the reference ships no golden vectors
 * of its own, so the goldens are minted from its running code.
 *
 * Third-party algorithms restated here (absent from the reference tree):
 *   numpy legacy RandomState (numpy>=1.17 keeps it frozen; verified against
 *   numpy 2.3.5): MT19937 init_genrand / genrand_int32, masked-rejection
 *   bounded integers (random_bounded_uint64_fill, use_masked=1),
 *   legacy_double (53 bit) and legacy_gauss (polar Box-Muller with cache).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MT_N 624
#define MT_M 397

typedef struct {
    uint32_t mt[MT_N];
    int pos;
    int has_gauss;
    double gauss;
} orc_rng;

/* numpy legacy seeding: np.random.seed(int) == mt19937 init_genrand(seed);
 * also clears the cached gaussian (custom_sampler.py:15, BPRMF_model.py:24). */
void orc_seed(orc_rng *r, uint32_t seed) {
    r->mt[0] = seed;
    for (int i = 1; i < MT_N; i++)
        r->mt[i] = 1812433253u * (r->mt[i - 1] ^ (r->mt[i - 1] >> 30)) + (uint32_t)i;
    r->pos = MT_N;
    r->has_gauss = 0;
    r->gauss = 0.0;
}

static void mt_twist(orc_rng *r) {
    uint32_t *mt = r->mt;
    int k;
    uint32_t y;
    for (k = 0; k < MT_N - MT_M; k++) {
        y = (mt[k] & 0x80000000u) | (mt[k + 1] & 0x7fffffffu);
        mt[k] = mt[k + MT_M] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    for (; k < MT_N - 1; k++) {
        y = (mt[k] & 0x80000000u) | (mt[k + 1] & 0x7fffffffu);
        mt[k] = mt[k + (MT_M - MT_N)] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    y = (mt[MT_N - 1] & 0x80000000u) | (mt[0] & 0x7fffffffu);
    mt[MT_N - 1] = mt[MT_M - 1] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    r->pos = 0;
}

uint32_t orc_next32(orc_rng *r) {
    if (r->pos == MT_N) mt_twist(r);
    uint32_t y = r->mt[r->pos++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

/* raw tempered stream, for pinning the device MT19937 generator */
void orc_raw_fill(orc_rng *r, uint32_t *out, int64_t n) {
    for (int64_t i = 0; i < n; i++) out[i] = orc_next32(r);
}

/* np.random.randint(n) (legacy, dtype=int64, low=0): rng = n-1; rng==0 returns
 * 0 WITHOUT consuming a draw; rng <= 0xffffffff: 32-bit masked rejection. */
int64_t orc_randint(orc_rng *r, int64_t n) {
    uint64_t rng = (uint64_t)(n - 1);
    if (rng == 0) return 0;
    uint32_t mask = (uint32_t)rng;
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4;
    mask |= mask >> 8; mask |= mask >> 16;
    uint32_t v;
    do { v = orc_next32(r) & mask; } while (v > (uint32_t)rng);
    return (int64_t)v;
}

static double orc_double(orc_rng *r) {
    uint32_t a = orc_next32(r) >> 5, b = orc_next32(r) >> 6;
    return (a * 67108864.0 + b) / 9007199254740992.0;
}

static double orc_gauss(orc_rng *r) {
    if (r->has_gauss) {
        r->has_gauss = 0;
        double t = r->gauss;
        r->gauss = 0.0;
        return t;
    }
    double f, x1, x2, r2;
    do {
        x1 = 2.0 * orc_double(r) - 1.0;
        x2 = 2.0 * orc_double(r) - 1.0;
        r2 = x1 * x1 + x2 * x2;
    } while (r2 >= 1.0 || r2 == 0.0);
    f = sqrt(-2.0 * log(r2) / r2);
    r->gauss = f * x1;
    r->has_gauss = 1;
    return f * x2;
}

/* np.random.normal(loc, scale, size) legacy: loc + scale * gauss, row-major. */
void orc_normal_fill(orc_rng *r, double loc, double scale, double *out, int64_t n) {
    for (int64_t i = 0; i < n; i++) out[i] = loc + scale * orc_gauss(r);
}

/* MFModel.__init__/initialize — BPRMF_model.py:15-56: seed, zero biases,
 * U ~ N(0,0.1) (n_users x d) drawn first, then V ~ N(0,0.1) (n_items x d). */
void orc_mf_init(uint32_t seed, int64_t n_users, int64_t n_items, int d,
                 double *U, double *V, double *item_bias) {
    orc_rng r;
    orc_seed(&r, seed);
    memset(item_bias, 0, sizeof(double) * (size_t)n_items);
    orc_normal_fill(&r, 0.0, 0.1, U, n_users * d);
    orc_normal_fill(&r, 0.0, 0.1, V, n_items * d);
}

/* Sampler.step / sample() — dataset/samplers/custom_sampler.py:24-46.
 * indptr/indices hold _ui_dict in the reference's list(set(...)) order
 * (custom_sampler.py:21); the host (Python) produces that order.
 * `u = randint(n_users)`; `i = ui[randint(lui)]`; `j = randint(n_items)`
 * redrawn while j in ui.  A user owning every item makes the reference
 * recurse and then spin forever (custom_sampler.py:35-41): returns -1 here.
 * Returns the number of raw 32-bit draws consumed (>=0) or -1. */
int64_t orc_sampler_step(orc_rng *r, int64_t n_users, int64_t n_items,
                         const int64_t *indptr, const int32_t *indices,
                         int64_t events, int32_t *out_u, int32_t *out_i, int32_t *out_j) {
    int64_t draws = 0;
    for (int64_t t = 0; t < events; t++) {
        /* count draws by re-implementing the masked rejection inline */
        int64_t u = -1, i = -1, j = -1;
        {
            uint64_t rng = (uint64_t)(n_users - 1);
            if (rng == 0) u = 0; else {
                uint32_t mask = (uint32_t)rng;
                mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
                uint32_t v; do { v = orc_next32(r) & mask; draws++; } while (v > (uint32_t)rng);
                u = v;
            }
        }
        const int32_t *ui = indices + indptr[u];
        int64_t lui = indptr[u + 1] - indptr[u];
        if (lui == n_items || lui == 0) return -1;
        {
            uint64_t rng = (uint64_t)(lui - 1);
            int64_t p;
            if (rng == 0) p = 0; else {
                uint32_t mask = (uint32_t)rng;
                mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
                uint32_t v; do { v = orc_next32(r) & mask; draws++; } while (v > (uint32_t)rng);
                p = v;
            }
            i = ui[p];
        }
        for (;;) {
            uint64_t rng = (uint64_t)(n_items - 1);
            if (rng == 0) j = 0; else {
                uint32_t mask = (uint32_t)rng;
                mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
                uint32_t v; do { v = orc_next32(r) & mask; draws++; } while (v > (uint32_t)rng);
                j = v;
            }
            int found = 0;
            for (int64_t q = 0; q < lui; q++) if (ui[q] == (int32_t)j) { found = 1; break; }
            if (!found) break;
        }
        out_u[t] = (int32_t)u; out_i[t] = (int32_t)i; out_j[t] = (int32_t)j;
    }
    return draws;
}

/* MFModel.update_factors — BPRMF_model.py:91-117 (train_step :87-89 loops it
 * one triple at a time).  float64, strictly sequential.  Aliasing quirk kept:
 * `user_factors` is a view, so after line 109 rewrites U[u] the item updates
 * (lines 112,116) read the UPDATED user row; V_i, V_j, biases and z are the
 * pre-update values.  Dot products are accumulated left to right (NumPy's
 * BLAS ddot order is unspecified; differences are O(1e-16) relative). */
void orc_bpr_update_seq(double *U, double *V, double *item_bias, int d,
                        double lr, double reg_u, double reg_b, double reg_pos, double reg_neg,
                        const int32_t *tu, const int32_t *ti, const int32_t *tj, int64_t n) {
    for (int64_t t = 0; t < n; t++) {
        double *u = U + (int64_t)tu[t] * d;
        double *vi = V + (int64_t)ti[t] * d;
        double *vj = V + (int64_t)tj[t] * d;
        double bi = item_bias[ti[t]], bj = item_bias[tj[t]];
        double xi = 0.0, xj = 0.0;
        for (int k = 0; k < d; k++) { xi += u[k] * vi[k]; xj += u[k] * vj[k]; }
        xi += bi; xj += bj;
        double z = 1.0 / (1.0 + exp(xi - xj));
        item_bias[ti[t]] = bi + lr * (z - reg_b * bi);
        item_bias[tj[t]] = bj + lr * (-z - reg_b * bj);
        for (int k = 0; k < d; k++) {
            double uk = u[k], ik = vi[k], jk = vj[k];
            double un = uk + lr * ((ik - jk) * z - reg_u * uk);
            u[k] = un;
            vi[k] = ik + lr * (un * z - reg_pos * ik);
            vj[k] = jk + lr * (-un * z - reg_neg * jk);
        }
    }
}

/* Sum over triples of softplus(-(x_ui - x_uj)) evaluated on the CURRENT
 * tables without updating: the BPR objective the kernels report. */
double orc_bpr_loss(const double *U, const double *V, const double *item_bias, int d,
                    const int32_t *tu, const int32_t *ti, const int32_t *tj, int64_t n) {
    double loss = 0.0;
    for (int64_t t = 0; t < n; t++) {
        const double *u = U + (int64_t)tu[t] * d;
        const double *vi = V + (int64_t)ti[t] * d;
        const double *vj = V + (int64_t)tj[t] * d;
        double x = item_bias[ti[t]] - item_bias[tj[t]];
        for (int k = 0; k < d; k++) x += u[k] * (vi[k] - vj[k]);
        loss += (x > 0) ? log1p(exp(-x)) : (-x + log1p(exp(x)));
    }
    return loss;
}

/* MFModel.get_user_predictions — BPRMF_model.py:70-85: b = item_bias + U[u]@V.T;
 * train items -> -inf; top-k by score descending.  Tie rule here: lower item
 * index first (the reference's argpartition/argsort[::-1] order for exact ties
 * is implementation-defined; TF path `tf.nn.top_k` is lower-index-first).
 * mask rows are the user's TRAIN items (CSR, any order).  Entries that are
 * -inf are never emitted: out_idx = -1, out_val = -inf pad the tail. */
void orc_user_topk(const double *U, const double *V, const double *item_bias,
                   int64_t n_items, int d,
                   const int64_t *mask_indptr, const int32_t *mask_indices,
                   const int32_t *users, int64_t n_sel, int k,
                   int32_t *out_idx, double *out_val) {
    double *s = (double *)malloc(sizeof(double) * (size_t)n_items);
    for (int64_t q = 0; q < n_sel; q++) {
        int64_t u = users[q];
        const double *ur = U + u * d;
        for (int64_t it = 0; it < n_items; it++) {
            const double *vr = V + it * d;
            double acc = 0.0;
            for (int c = 0; c < d; c++) acc += ur[c] * vr[c];
            s[it] = (item_bias ? item_bias[it] : 0.0) + acc;
        }
        if (mask_indptr)
            for (int64_t p = mask_indptr[u]; p < mask_indptr[u + 1]; p++) s[mask_indices[p]] = -INFINITY;
        for (int r = 0; r < k; r++) {
            int64_t best = -1;
            double bv = -INFINITY;
            for (int64_t it = 0; it < n_items; it++)
                if (s[it] > bv) { bv = s[it]; best = it; }
            out_idx[q * k + r] = (int32_t)best;
            out_val[q * k + r] = bv;
            if (best >= 0) s[best] = -INFINITY;
        }
    }
    free(s);
}

size_t orc_rng_size(void) { return sizeof(orc_rng); }

/* ---------------------------------------------------------------------------
 * MF2020 (pointwise logistic MF of "NCF vs MF revisited").
 * MFModel.train_step — elliot/recommender/latent_factor_models/MF2020/MF_model.py:80-112:
 *   pred = gb + ub[u] + ib[i] + U[u].V[i];  numerically split sigmoid/loss (:94-101);
 *   grad = rating - sigmoid;
 *   U[u] += lr (grad V[i] - reg U[u])          (uf_ is a VIEW: updated in place, :105)
 *   V[i] += lr (grad U'[u] - reg V[i])         (reads the UPDATED user row, :106)
 *   ub[u] += lr (grad - reg ub_old); ib[i] += lr (grad - reg ib_old); gb += lr (grad - reg gb_old)
 * batch_loss[t / batch] accumulates this_loss in sample order (MF.py:123: loss += train_step(batch)/len(batch)
 * is left to the caller).  Strictly sequential: every sample reads and writes the global bias. */
void orc_mf2020_update_seq(double *U, double *V, double *ub, double *ib, double *gb, int d, double lr, double reg,
                           const int32_t *su, const int32_t *si, const int32_t *sr, int64_t n, int64_t batch,
                           double *batch_loss) {
    double g = *gb;
    for (int64_t t = 0; t < n; t++) {
        double *u = U + (int64_t)su[t] * d;
        double *v = V + (int64_t)si[t] * d;
        const double r = (double)sr[t];
        const double bu = ub[su[t]], bi = ib[si[t]];
        double dot = 0.0;
        for (int k = 0; k < d; k++) dot += u[k] * v[k];
        const double pred = g + bu + bi + dot;
        double sig, loss;
        if (pred > 0) {
            const double ope = 1.0 + exp(-pred);
            sig = 1.0 / ope;
            loss = log(ope) + (1.0 - r) * pred;
        } else {
            const double e = exp(pred);
            sig = e / (1.0 + e);
            loss = -r * pred + log(1.0 + e);
        }
        const double grad = r - sig;
        for (int k = 0; k < d; k++) {
            const double uk = u[k], vk = v[k];
            const double un = uk + lr * (grad * vk - reg * uk);
            u[k] = un;
            v[k] = vk + lr * (grad * un - reg * vk);
        }
        ub[su[t]] = bu + lr * (grad - reg * bu);
        ib[si[t]] = bi + lr * (grad - reg * bi);
        g += lr * (grad - reg * g);
        if (batch_loss) batch_loss[t / batch] += loss;
    }
    *gb = g;
}
