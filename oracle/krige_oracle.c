/*
 * oracle/krige_oracle.c -- CPU restatement of the GSTools kriging evaluation
 * (SURVEY.md section 8f, row f1).  TEST INFRASTRUCTURE ONLY: only tests/,
 * __graft_entry__.smoke() and bench.py's CPU legs may load this library.
 *
 * What it restates.  `calc_field_krige_and_variance` / `calc_field_krige` live in the external
 * package gstools-cython (>=1,<2, /root/reference/pyproject.toml:44; Rust twin in gstools_core),
 * not vendored under /root/reference.  The reference imports them at
 * src/gstools/krige/base.py:16-19 (Cython) / :30-33 (Rust) and calls them through
 * `_calc_field_krige[_and_variance]` (base.py:42-61) from `Krige._summate` (base.py:307-317) with
 *     krig_mat  (K, K)  the (pseudo-)inverted kriging matrix        base.py:319-357
 *     krig_vecs (K, n)  right-hand sides of one chunk of points     base.py:359-388
 *     cond      (K,)    conditioning values, zero padded            base.py:562-565
 * and expects  field[k] = sum_i cond[i] (M kv)[i,k],  error[k] = sum_i kv[i,k] (M kv)[i,k];
 * the caller then forms sill - error (base.py:296-298).
 *
 * Published loop nest (gstools-cython krige.pyx): points outer (parallel, one owner per point),
 * for each matrix row i the dot product over j ascending, then the two accumulations; plain fp64.
 *
 * Parity pin.  tests/test_oracle_golden.py: with this library bound as the reference's
 * gstools_cython.krige, the known-answer checks of the reference's tests/test_krige.py (kriged
 * field reproduces the conditioning values, error variance equals the nugget there,
 * structured == unstructured) hold; boundary arrays recorded by tests/golden/make_golden_krige.py.
 */
#include <stddef.h>
#include <stdint.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static int pick_threads(int num_threads)
{
#ifdef _OPENMP
    if (num_threads <= 0) return omp_get_max_threads();
    return num_threads;
#else
    (void)num_threads;
    return 1;
#endif
}

/* krig_mat (K,K) row-major, krig_vecs (K,n) row stride ld, cond (K,), field/error (n,) */
int oracle_calc_field_krige_and_variance(const double *krig_mat, const double *krig_vecs,
                                         int64_t ld, const double *cond, int64_t K, int64_t n,
                                         double *field, double *error, int num_threads)
{
    const int nt = pick_threads(num_threads);
    (void)nt;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(nt)
#endif
    for (int64_t k = 0; k < n; ++k) {
        double f = 0.0, e = 0.0;
        for (int64_t i = 0; i < K; ++i) {
            double krig_fac = 0.0;
            for (int64_t j = 0; j < K; ++j) krig_fac += krig_mat[i * K + j] * krig_vecs[j * ld + k];
            e += krig_vecs[i * ld + k] * krig_fac;
            f += cond[i] * krig_fac;
        }
        field[k] = f;
        error[k] = e;
    }
    return 0;
}

int oracle_calc_field_krige(const double *krig_mat, const double *krig_vecs, int64_t ld,
                            const double *cond, int64_t K, int64_t n, double *field,
                            int num_threads)
{
    const int nt = pick_threads(num_threads);
    (void)nt;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(nt)
#endif
    for (int64_t k = 0; k < n; ++k) {
        double f = 0.0;
        for (int64_t i = 0; i < K; ++i) {
            double krig_fac = 0.0;
            for (int64_t j = 0; j < K; ++j) krig_fac += krig_mat[i * K + j] * krig_vecs[j * ld + k];
            f += cond[i] * krig_fac;
        }
        field[k] = f;
    }
    return 0;
}
