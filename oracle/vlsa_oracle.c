/* TEST INFRASTRUCTURE — plain-C (fp64) restatement of the VLSA aggregation path.
 *
 * Independent of torch: a second opinion for the torch-based oracle (oracle/vlsa_oracle.py) and for the
 * CUDA kernels.  Parity status: PINNED — tests/test_oracle_c.py checks it against the golden vectors that
 * tests/golden/make_golden.py produced by running the unmodified reference.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library.
 *
 * Follows (liupei101/VLSA @ 915f37a):
 *   model/deepmil.py:187-204  VLFAN.forward         model/vlsa.py:185-192   cosine head
 *   utils/func.py:44          softmax converter     loss/loss_surv.py:153-164  SurvIFMLE
 *   loss/loss_surv_ext.py:42-55,81-102  SurvEMD     model/deepmil.py:16-37  logit_pooling
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define D 512
#define EPS_NORM 1e-12

static double dot(const double* a, const double* b, int n) { double s = 0; for (int i = 0; i < n; ++i) s += a[i] * b[i]; return s; }
static void normalize(const double* x, double* y, int n) {            /* F.normalize, eps 1e-12 */
    double nrm = sqrt(dot(x, x, n)); if (nrm < EPS_NORM) nrm = EPS_NORM;
    for (int i = 0; i < n; ++i) y[i] = x[i] / nrm;
}

/* X [N,D] float, Q [P,D], W [D,D], b [D], T [R,D]; scale = exp(fp32(log 100)); logit_scale = log-space parameter.
 * Outputs (double): f [D], g [D], logits [R], inc [R]; A [P,N] optional (may be NULL). Returns 0. */
int vlsa_oracle_forward(const float* X, int64_t N, const float* Q, int P, const float* W, const float* b,
                        const float* T, int R, double scale, double logit_scale,
                        double* f, double* g, double* logits, double* inc, double* A) {
    double* Qn = malloc(sizeof(double) * P * D);
    double* S = malloc(sizeof(double) * (size_t)P * (N > 0 ? N : 1));
    double* O = calloc((size_t)P * D, sizeof(double));
    double q[D], x[D], xn[D];
    if (!Qn || !S || !O) return -1;
    for (int p = 0; p < P; ++p) { for (int d = 0; d < D; ++d) q[d] = Q[p * D + d]; normalize(q, Qn + p * D, D); }   /* :187 */
    for (int64_t n = 0; n < N; ++n) {
        for (int d = 0; d < D; ++d) x[d] = X[n * D + d];
        normalize(x, xn, D);                                                                                  /* :189 */
        for (int p = 0; p < P; ++p) S[(size_t)p * N + n] = scale * dot(Qn + p * D, xn, D);                     /* :190,197 */
    }
    for (int p = 0; p < P; ++p) {                                                                             /* :198,200 */
        double m = -INFINITY, l = 0;
        for (int64_t n = 0; n < N; ++n) if (S[(size_t)p * N + n] > m) m = S[(size_t)p * N + n];
        for (int64_t n = 0; n < N; ++n) l += exp(S[(size_t)p * N + n] - m);
        for (int64_t n = 0; n < N; ++n) {
            const double a = exp(S[(size_t)p * N + n] - m) / l;
            if (A) A[(size_t)p * N + n] = a;
            for (int d = 0; d < D; ++d) O[p * D + d] += a * X[n * D + d];
        }
    }
    double v[D];
    for (int d = 0; d < D; ++d) { double s = 0; for (int p = 0; p < P; ++p) s += O[p * D + d]; v[d] = s / P; }   /* :136 */
    for (int o = 0; o < D; ++o) { double s = b[o]; for (int i = 0; i < D; ++i) s += (double)W[o * D + i] * v[i]; f[o] = s; }  /* :204 */
    normalize(f, g, D);                                                                                       /* vlsa.py:189 */
    const double ls = exp(logit_scale);
    double tn[D], t[D], mx = -INFINITY, den = 0;
    for (int r = 0; r < R; ++r) {
        for (int d = 0; d < D; ++d) t[d] = T[r * D + d];
        normalize(t, tn, D);                                                                                  /* vlsa.py:186 */
        double s = 0; for (int d = 0; d < D; ++d) s += (ls * g[d]) * tn[d];                                    /* vlsa.py:192 */
        logits[r] = s; if (s > mx) mx = s;
    }
    for (int r = 0; r < R; ++r) den += exp(logits[r] - mx);
    for (int r = 0; r < R; ++r) inc[r] = exp(logits[r] - mx) / den;                                           /* func.py:44 */
    free(Qn); free(S); free(O);
    return 0;
}

/* p [B,R] incidence (already softmaxed), t/e [B]; out[0] = mean SurvIFMLE, out[1] = mean SurvEMD (p=2, raw). */
int vlsa_oracle_losses(const double* p, const int64_t* t, const int64_t* e, int B, int R, double ls_exp,
                       double alpha, double eps, double* out) {
    double s1 = 0, s2 = 0;
    for (int i = 0; i < B; ++i) {
        const double* pi = p + (size_t)i * R;
        const int ti = (int)t[i]; const double c = 1.0 - (double)e[i];
        double cif = 0; for (int r = 0; r <= ti; ++r) cif += pi[r];
        const double pt = pi[ti] > eps ? pi[ti] : eps, st = (1 - cif) > eps ? (1 - cif) : eps;
        const double unc = -(1 - c) * log(pt), cen = -c * log(st);
        s1 += (1 - alpha) * (cen + unc) + alpha * unc;                                      /* loss_surv.py:157-161 */
        double tl[64], pr[64], m1 = -INFINITY, m2 = -INFINITY, d1 = 0, d2 = 0;
        for (int r = 0; r < R; ++r) {
            const double target = (r == ti) ? 1.0 : (r > ti ? (1.0 - (double)e[i]) : 0.0);   /* loss_surv_ext.py:42-55 */
            tl[r] = (2 * target - 1) * ls_exp;                                              /* :90 */
            pr[r] = (1 - (double)e[i]) * ((1 - target) * pi[r] + target * ls_exp) + (double)e[i] * pi[r];  /* :93 */
            if (tl[r] > m1) m1 = tl[r]; if (pr[r] > m2) m2 = pr[r];
        }
        for (int r = 0; r < R; ++r) { d1 += exp(tl[r] - m1); d2 += exp(pr[r] - m2); }
        double ct = 0, cp = 0, acc = 0;
        for (int r = 0; r < R; ++r) {
            ct += exp(tl[r] - m1) / d1; cp += exp(pr[r] - m2) / d2;
            acc += (cp - ct) * (cp - ct);                                                   /* :13-40, p=2 raw */
        }
        s2 += acc;
    }
    out[0] = s1 / B; out[1] = s2 / B;
    return 0;
}

static int cmp_desc(const void* a, const void* b) { const double x = *(const double*)a, y = *(const double*)b; return (x < y) - (x > y); }

/* zero-shot arm: per-patch logits ls * cos(x_n, T_r), pooled per class; mode 0 mean, 1 top-k mean. pred = first argmax. */
int vlsa_oracle_logit_pool(const float* X, int64_t N, const float* T, int R, double logit_scale, int mode, int k,
                           double* pooled, int64_t* pred) {
    double* L = malloc(sizeof(double) * (size_t)N * R);
    double* Tn = malloc(sizeof(double) * R * D);
    double x[D], xn[D], t[D];
    if (!L || !Tn) return -1;
    const double ls = exp(logit_scale);
    for (int r = 0; r < R; ++r) { for (int d = 0; d < D; ++d) t[d] = T[r * D + d]; normalize(t, Tn + r * D, D); }
    for (int64_t n = 0; n < N; ++n) {
        for (int d = 0; d < D; ++d) x[d] = X[n * D + d];
        normalize(x, xn, D);
        for (int r = 0; r < R; ++r) { double s = 0; for (int d = 0; d < D; ++d) s += (ls * xn[d]) * Tn[r * D + d]; L[(size_t)r * N + n] = s; }
    }
    for (int r = 0; r < R; ++r) {
        double s = 0;
        if (mode == 0) { for (int64_t n = 0; n < N; ++n) s += L[(size_t)r * N + n]; pooled[r] = s / (double)N; }
        else {
            const int64_t kk = k < N ? k : N;
            qsort(L + (size_t)r * N, (size_t)N, sizeof(double), cmp_desc);
            for (int64_t j = 0; j < kk; ++j) s += L[(size_t)r * N + j];
            pooled[r] = s / (double)kk;
        }
    }
    int best = 0; for (int r = 1; r < R; ++r) if (pooled[r] > pooled[best]) best = r;
    *pred = best;
    free(L); free(Tn);
    return 0;
}
