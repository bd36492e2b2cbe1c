/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY.  Never imported, linked or executed by the product path
 * (transformer-transducer_b200/, warprnnt_pytorch/).  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may use it, and only as the checker / CPU baseline.
 *
 * Plain-C restatement of the CPU transducer-loss algorithm that the reference calls through the
 * third-party module `warprnnt_pytorch` (HawkAaron/warp-transducer, un-vendored and un-pinned:
 * /root/reference/requirements.txt:6; imported /root/reference/train.py:13, constructed train.py:231,
 * called train.py:53 and /root/reference/espnet/nets/pytorch_backend/transducer/loss.py:23-25,74).
 * The package source is not on disk, so this file restates its published CPU algorithm
 * (upstream include/detail/cpu_rnnt.h): inputs are LOG-PROBABILITIES (the Python binding applies
 * torch.log_softmax before calling the CPU kernel), per utterance it runs the alpha recursion, the
 * beta recursion, and writes the sparse gradient w.r.t. the log-probabilities; autograd then
 * chains through log_softmax.  Semantics are spelled out in SURVEY.md section 8(a) row a6.
 *
 * Parity pin: upstream known-answer vector (cost 4.495666, tests/golden/warp_transducer_kat.json)
 * and torchaudio.functional.rnnt_loss cross-check (tests/test_oracle.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline double lse2(double a, double b) {
    if (a == -INFINITY) return b;
    if (b == -INFINITY) return a;
    return a > b ? a + log1p(exp(b - a)) : b + log1p(exp(a - b));
}

static inline float lse2f(float a, float b) {
    if (a == -INFINITY) return b;
    if (b == -INFINITY) return a;
    return a > b ? a + log1pf(expf(b - a)) : b + log1pf(expf(a - b));
}

/* One utterance.  lp: (maxT, maxU1, V) log-probs of this utterance; grad: same shape, pre-zeroed;
 * T, U1 = U_b + 1 the valid extents.  Instantiated for float (upstream's ProbT) and double (arbiter). */
#define DEFINE_UTT(NAME, REAL, LSE2, EXP)                                                              \
    static REAL NAME(const REAL* lp, REAL* grad, const int32_t* labels, int T, int U1, int maxU1, int V, \
                     int blank, REAL* alpha, REAL* beta) {                                             \
        AL(0, 0) = 0;                                                                                  \
        for (int t = 0; t < T; ++t)                                                                    \
            for (int u = 0; u < U1; ++u) {                                                             \
                if (u == 0 && t > 0) AL(t, 0) = AL(t - 1, 0) + LP(t - 1, 0, blank);                    \
                if (t == 0 && u > 0) AL(0, u) = AL(0, u - 1) + LP(0, u - 1, labels[u - 1]);            \
                if (t > 0 && u > 0) {                                                                  \
                    REAL no_emit = AL(t - 1, u) + LP(t - 1, u, blank);                                 \
                    REAL emit = AL(t, u - 1) + LP(t, u - 1, labels[u - 1]);                            \
                    AL(t, u) = LSE2(emit, no_emit);                                                    \
                }                                                                                      \
            }                                                                                          \
        REAL ll_fwd = AL(T - 1, U1 - 1) + LP(T - 1, U1 - 1, blank);                                    \
        BE(T - 1, U1 - 1) = LP(T - 1, U1 - 1, blank);                                                  \
        for (int t = T - 1; t >= 0; --t)                                                               \
            for (int u = U1 - 1; u >= 0; --u) {                                                        \
                if (u == U1 - 1 && t < T - 1) BE(t, U1 - 1) = BE(t + 1, U1 - 1) + LP(t, U1 - 1, blank); \
                if (t == T - 1 && u < U1 - 1) BE(T - 1, u) = BE(T - 1, u + 1) + LP(T - 1, u, labels[u]); \
                if (t < T - 1 && u < U1 - 1) {                                                         \
                    REAL no_emit = BE(t + 1, u) + LP(t, u, blank);                                     \
                    REAL emit = BE(t, u + 1) + LP(t, u, labels[u]);                                    \
                    BE(t, u) = LSE2(emit, no_emit);                                                    \
                }                                                                                      \
            }                                                                                          \
        REAL ll = BE(0, 0);                                                                            \
        if (grad) {                                                                                    \
            for (int t = 0; t < T; ++t)                                                                \
                for (int u = 0; u < U1; ++u) {                                                         \
                    if (t < T - 1) GR(t, u, blank) = -EXP(LP(t, u, blank) + AL(t, u) + BE(t + 1, u) - ll); \
                    if (u < U1 - 1)                                                                    \
                        GR(t, u, labels[u]) = -EXP(LP(t, u, labels[u]) + AL(t, u) + BE(t, u + 1) - ll); \
                }                                                                                      \
            GR(T - 1, U1 - 1, blank) = -EXP(LP(T - 1, U1 - 1, blank) + AL(T - 1, U1 - 1) - ll);        \
        }                                                                                              \
        return -ll_fwd;                                                                                \
    }
#define LP(t, u, v) lp[((size_t)(t) * maxU1 + (u)) * V + (v)]
#define GR(t, u, v) grad[((size_t)(t) * maxU1 + (u)) * V + (v)]
#define AL(t, u) alpha[(size_t)(t) * U1 + (u)]
#define BE(t, u) beta[(size_t)(t) * U1 + (u)]
DEFINE_UTT(utt_f32, float, lse2f, expf)
DEFINE_UTT(utt_f64, double, lse2, exp)
#undef LP
#undef GR
#undef AL
#undef BE

/* Batch entry (float32).  log_probs, grads: (B, maxT, maxU1, V) contiguous; labels: (B, maxU1-1)
 * int32 (entries at u >= label_lens[b] are never read, they may be -1: /root/reference/tt/dataset.py:46-48);
 * costs: (B).  grads may be NULL.  Returns 0, or the 1-based index of the first bad utterance. */
int oracle_rnnt_f32(const float* log_probs, const int32_t* labels, const int32_t* act_lens,
                    const int32_t* label_lens, int B, int maxT, int maxU1, int V, int blank,
                    float* costs, float* grads) {
    int bad = 0;
    if (grads) memset(grads, 0, (size_t)B * maxT * maxU1 * V * sizeof(float));
#pragma omp parallel for schedule(dynamic)
    for (int b = 0; b < B; ++b) {
        int T = act_lens[b], U1 = label_lens[b] + 1;
        if (T < 1 || T > maxT || U1 < 1 || U1 > maxU1) {
#pragma omp critical
            if (!bad) bad = b + 1;
            continue;
        }
        float* alpha = (float*)malloc(sizeof(float) * 2 * (size_t)T * U1);
        float* beta = alpha + (size_t)T * U1;
        size_t off = (size_t)b * maxT * maxU1 * V;
        costs[b] = utt_f32(log_probs + off, grads ? grads + off : NULL, labels + (size_t)b * (maxU1 - 1),
                           T, U1, maxU1, V, blank, alpha, beta);
        free(alpha);
    }
    return bad;
}

/* Same batch entry in float64 (arbiter for lattices long enough that float32 alpha/beta rounding shows). */
int oracle_rnnt_f64(const double* log_probs, const int32_t* labels, const int32_t* act_lens,
                    const int32_t* label_lens, int B, int maxT, int maxU1, int V, int blank,
                    double* costs, double* grads) {
    int bad = 0;
    if (grads) memset(grads, 0, (size_t)B * maxT * maxU1 * V * sizeof(double));
#pragma omp parallel for schedule(dynamic)
    for (int b = 0; b < B; ++b) {
        int T = act_lens[b], U1 = label_lens[b] + 1;
        if (T < 1 || T > maxT || U1 < 1 || U1 > maxU1) {
#pragma omp critical
            if (!bad) bad = b + 1;
            continue;
        }
        double* alpha = (double*)malloc(sizeof(double) * 2 * (size_t)T * U1);
        double* beta = alpha + (size_t)T * U1;
        size_t off = (size_t)b * maxT * maxU1 * V;
        costs[b] = utt_f64(log_probs + off, grads ? grads + off : NULL, labels + (size_t)b * (maxU1 - 1),
                           T, U1, maxU1, V, blank, alpha, beta);
        free(alpha);
    }
    return bad;
}

/* float64 arbiter on GATHERED log-probs: lpb/lpl (B, maxT, maxU1) = log p(blank), log p(label_{u+1}).
 * Writes alpha/beta (B, maxT, maxU1) (untouched outside the valid region) and costs. */
int oracle_lattice_f64(const double* lpb, const double* lpl, const int32_t* act_lens,
                       const int32_t* label_lens, int B, int maxT, int maxU1, double* alpha,
                       double* beta, double* costs) {
    for (int b = 0; b < B; ++b) {
        int T = act_lens[b], U1 = label_lens[b] + 1;
        if (T < 1 || T > maxT || U1 < 1 || U1 > maxU1) return b + 1;
        size_t o = (size_t)b * maxT * maxU1;
#define I(t, u) (o + (size_t)(t) * maxU1 + (u))
        for (int t = 0; t < T; ++t)
            for (int u = 0; u < U1; ++u) {
                double v;
                if (t == 0 && u == 0) v = 0.0;
                else if (u == 0) v = alpha[I(t - 1, 0)] + lpb[I(t - 1, 0)];
                else if (t == 0) v = alpha[I(0, u - 1)] + lpl[I(0, u - 1)];
                else v = lse2(alpha[I(t - 1, u)] + lpb[I(t - 1, u)], alpha[I(t, u - 1)] + lpl[I(t, u - 1)]);
                alpha[I(t, u)] = v;
            }
        for (int t = T - 1; t >= 0; --t)
            for (int u = U1 - 1; u >= 0; --u) {
                double v;
                if (t == T - 1 && u == U1 - 1) v = lpb[I(t, u)];
                else if (u == U1 - 1) v = beta[I(t + 1, u)] + lpb[I(t, u)];
                else if (t == T - 1) v = beta[I(t, u + 1)] + lpl[I(t, u)];
                else v = lse2(beta[I(t + 1, u)] + lpb[I(t, u)], beta[I(t, u + 1)] + lpl[I(t, u)]);
                beta[I(t, u)] = v;
            }
        costs[b] = -(alpha[I(T - 1, U1 - 1)] + lpb[I(T - 1, U1 - 1)]);
#undef I
    }
    return 0;
}
