/*
 * TEST INFRASTRUCTURE - CPU oracle for the WGSassign genotype-likelihood hot path.
 *
 * A plain-C restatement of the arithmetic of the reference's Cython kernels, used ONLY
 * by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * as the checker.  Nothing under wgsassign_b200/ may link, import or call this file.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks every function here bit-for-bit
 * against (a) the reference's own compiled kernels (oracle/_ref, built from
 * /root/reference by oracle/build_ref.sh) on seeded inputs and (b) the golden fixtures
 * in tests/golden/ that were produced by running the unmodified reference CLI.
 *
 * Rounding model (checked in the C that Cython 3.3 generates from the reference .pyx):
 * locals are `float`; an integer literal inside a float expression is emitted as a
 * DOUBLE literal (1.0 / 2.0), so a sub-expression that touches a literal is evaluated
 * in double while float(op)float sub-expressions stay float; the statement's value is
 * narrowed to float once, on assignment.  Each expression below therefore keeps the
 * reference's operand order and parenthesisation, with `1.0`/`2.0` double literals
 * exactly where the generated C has them.  Build WITHOUT -ffast-math / -ffp-contract.
 *
 * Sites are independent in every kernel, so the OpenMP `parallel for` over sites (the
 * reference's `prange`) cannot change any result.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ROW(p, s, ld) ((p) + (size_t)(s) * (size_t)(ld))

/* emMAF_cy.pyx:10-23 (emMAF_update): one EM step of f over the n individuals of L.
 * L is [m, 2n] (g0,g1 interleaved), f is [m], updated in place. */
void orc_em_update(const float *L, int m, int n, float *f, int t)
{
    int s;
#pragma omp parallel for num_threads(t) schedule(static)
    for (s = 0; s < m; ++s) {
        const float *row = ROW(L, s, 2 * n);
        float fs = f[s];
        float tmp = 0.0;
        for (int i = 0; i < n; ++i) {
            float l0 = row[2 * i + 0], l1 = row[2 * i + 1];
            float p0 = (l0 * (1.0 - fs)) * (1.0 - fs);          /* :19 */
            float p1 = ((l1 * 2.0) * fs) * (1.0 - fs);          /* :20 */
            float p2 = (((1.0 - l0) - l1) * fs) * fs;           /* :21 */
            tmp = tmp + ((p1 + (2.0 * p2)) / (2.0 * ((p0 + p1) + p2))); /* :22 */
        }
        f[s] = tmp / ((float)n);                                /* :23 */
    }
}

/* emMAF_cy.pyx:26-33 (rmse1d): serial float accumulation, float divide, double sqrt. */
double orc_rmse1d(const float *v1, const float *v2, int n)
{
    float res = 0.0;
    for (int i = 0; i < n; ++i)
        res = res + ((v1[i] - v2[i]) * (v1[i] - v2[i]));
    res = res / ((float)n);
    return sqrt(res);
}

/* emMAF.py:15-27 (emMAF): f = 0.25, up to `iter` updates, stop when rmse < tole AFTER
 * the update.  Returns the 1-based iteration it converged at, 0 if it never did.
 * `prev` is caller scratch of m floats. */
int orc_emMAF(const float *L, int m, int n, int iter, double tole, float *f, float *prev, int t)
{
    for (int s = 0; s < m; ++s) { f[s] = 0.25f; prev[s] = 0.25f; }
    for (int it = 0; it < iter; ++it) {
        orc_em_update(L, m, n, f, t);
        double diff = orc_rmse1d(f, prev, m);
        if (diff < tole) return it + 1;
        for (int s = 0; s < m; ++s) prev[s] = f[s];
    }
    return 0;
}

/* glassy_cy.pyx:12-21 (loglike): vec[s] += log(GL . HWE(A[s,k])) for individual i.
 * L is [m, ldl] with ldl = 2N; A is [m, K]. */
void orc_loglike(const float *L, int m, int ldl, const float *A, int K, int i, int k,
                 float *vec, int t)
{
    int s;
#pragma omp parallel for num_threads(t) schedule(static)
    for (s = 0; s < m; ++s) {
        float l0 = ROW(L, s, ldl)[2 * i + 0], l1 = ROW(L, s, ldl)[2 * i + 1];
        float a = ROW(A, s, K)[k];
        float like0 = (l0 * (1.0 - a)) * (1.0 - a);             /* :18 */
        float like1 = ((l1 * 2.0) * (1.0 - a)) * a;             /* :19 */
        float like2 = (((1.0 - l0) - l1) * a) * a;              /* :20 */
        vec[s] = vec[s] + log((like0 + like1) + like2);         /* :21 */
    }
}

/* fisher_cy.pyx:12-30 (fisher_obs): f_pop[s] += sum_r term(L[s,r], A[s,i]); L is the
 * population's own [m, 2n] column copy, A the full [m, K] matrix. */
static inline float fisher_term(float g0, float g1, float th)
{
    float g2 = (1.0 - g0) - g1;                                              /* :24 */
    float u = (((g0 * (1.0 - th)) * (1.0 - th)) + (((g1 * 2.0) * th) * (1.0 - th)))
              + ((g2 * th) * th);                                            /* :25 */
    float n1 = 2.0 * ((g0 + g2) - (2.0 * g1));                               /* :26 */
    float n2 = (th * n1) + (2.0 * (g1 - g0));                                /* :27 */
    float term = -1.0 * ((n1 / u) - ((n2 / u) * (n2 / u)));                  /* :28 */
    return term;
}

void orc_fisher_obs(const float *L, int m, int n, const float *A, int K, int i, float *f_pop, int t)
{
    int s;
#pragma omp parallel for num_threads(t) schedule(static)
    for (s = 0; s < m; ++s) {
        const float *row = ROW(L, s, 2 * n);
        float term_sum = 0;
        float th = ROW(A, s, K)[i];
        for (int r = 0; r < n; ++r) {
            float term = fisher_term(row[2 * r + 0], row[2 * r + 1], th);
            term_sum = term_sum + term;                                      /* :29 */
        }
        f_pop[s] = f_pop[s] + term_sum;                                      /* :30 */
    }
}

/* fisher_cy.pyx:32-39 (ne_obs) and :58-65 (ne_obs_ind): out[s] += 0.5*f[s]*A*(1-A). */
void orc_ne_obs(const float *f_pop, int m, const float *A, int K, int i, float *ne_pop, int t)
{
    int s;
#pragma omp parallel for num_threads(t) schedule(static)
    for (s = 0; s < m; ++s) {
        float a = ROW(A, s, K)[i];
        float n_tilde = ((0.5 * f_pop[s]) * a) * (1.0 - a);                  /* :38 */
        ne_pop[s] = ne_pop[s] + n_tilde;
    }
}

/* fisher_cy.pyx:41-56 (fisher_obs_ind): one individual i of the full [m, ldl] matrix. */
void orc_fisher_obs_ind(const float *L, int m, int ldl, const float *A, int K, int i, int pop_i,
                        float *f_ind, int t)
{
    int s;
#pragma omp parallel for num_threads(t) schedule(static)
    for (s = 0; s < m; ++s) {
        float th = ROW(A, s, K)[pop_i];
        float term = fisher_term(ROW(L, s, ldl)[2 * i + 0], ROW(L, s, ldl)[2 * i + 1], th);
        f_ind[s] = f_ind[s] + term;                                          /* :56 */
    }
}

/* zscore_cy.pyx:10-34 (expected_W_l).  A is the AF VECTOR over kept sites (indexed by
 * position in L_keep, not by site).  AD is [m, ldl] int32, AD_factorial/AD_like [C,3],
 * AD_index [idx_rows, idx_cols] read as AD_index[Aa, Ar] (zscore_cy.pyx:30). */
void orc_expected_W_l(const float *L, int ldl, const int *L_keep, int mk, const float *A,
                      const int *AD, const float *AD_factorial, const float *AD_like,
                      const int *AD_index, int idx_cols, int i,
                      float *W_l_obs_array, float *W_l_array, int t)
{
    int si;
#pragma omp parallel for num_threads(t) schedule(static)
    for (si = 0; si < mk; ++si) {
        int s = L_keep[si];
        float A_sk = A[si];
        float P_gl0 = (1.0 - A_sk) * (1.0 - A_sk);                           /* :19 */
        float P_gl1 = (2.0 * (1.0 - A_sk)) * A_sk;                           /* :20 */
        float P_gl2 = A_sk * A_sk;                                           /* :21 */
        float l0 = ROW(L, s, ldl)[2 * i + 0], l1 = ROW(L, s, ldl)[2 * i + 1];
        float f_gl0 = l0 * P_gl0;
        float f_gl1 = l1 * P_gl1;
        float f_gl2 = ((1.0 - l0) - l1) * P_gl2;                             /* :24 */
        float f_gl_log = log((f_gl0 + f_gl1) + f_gl2);                       /* :25 */
        W_l_obs_array[si] = W_l_obs_array[si] + f_gl_log;
        int Dl = ROW(AD, s, ldl)[2 * i] + ROW(AD, s, ldl)[2 * i + 1];
        for (int Aa = 0; Aa < Dl + 1; ++Aa) {
            int Ar = Dl - Aa;
            int c = AD_index[(size_t)Aa * idx_cols + Ar];
            const float *lk = AD_like + 3 * (size_t)c, *fa = AD_factorial + 3 * (size_t)c;
            float e = log(((lk[0] * P_gl0) + (lk[1] * P_gl1)) + (lk[2] * P_gl2)); /* :31 */
            W_l_array[si] = W_l_array[si] + (((e * P_gl0) * fa[0]) * 1.0);   /* :32 */
            W_l_array[si] = W_l_array[si] + (((e * P_gl1) * fa[1]) * 1.0);   /* :33 */
            W_l_array[si] = W_l_array[si] + (((e * P_gl2) * fa[2]) * 1.0);   /* :34 */
        }
    }
}

/* zscore_cy.pyx:37-56 (variance_W_l). */
void orc_variance_W_l(const float *L, int ldl, const int *L_keep, int mk, const float *A,
                      const int *AD, const float *AD_factorial, const float *AD_like,
                      const int *AD_index, int idx_cols, int i,
                      float *var_W_l_array, const float *W_l_array, int t)
{
    (void)L;
    int si;
#pragma omp parallel for num_threads(t) schedule(static)
    for (si = 0; si < mk; ++si) {
        int s = L_keep[si];
        float A_sk = A[si];
        float P_gl0 = (1.0 - A_sk) * (1.0 - A_sk);
        float P_gl1 = (2.0 * (1.0 - A_sk)) * A_sk;
        float P_gl2 = A_sk * A_sk;
        int Dl = ROW(AD, s, ldl)[2 * i] + ROW(AD, s, ldl)[2 * i + 1];
        for (int Aa = 0; Aa < Dl + 1; ++Aa) {
            int Ar = Dl - Aa;
            int c = AD_index[(size_t)Aa * idx_cols + Ar];
            const float *lk = AD_like + 3 * (size_t)c, *fa = AD_factorial + 3 * (size_t)c;
            float e = log(((lk[0] * P_gl0) + (lk[1] * P_gl1)) + (lk[2] * P_gl2));
            var_W_l_array[si] = var_W_l_array[si]
                + (((powf(W_l_array[si] - e, 2.0) * P_gl0) * fa[0]) * 1.0);  /* :54 */
            var_W_l_array[si] = var_W_l_array[si]
                + (((powf(W_l_array[si] - e, 2.0) * P_gl1) * fa[1]) * 1.0);  /* :55 */
            var_W_l_array[si] = var_W_l_array[si]
                + (((powf(W_l_array[si] - e, 2.0) * P_gl2) * fa[2]) * 1.0);  /* :56 */
        }
    }
}

int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
