/*
 * oracle/summate_oracle.c -- CPU restatement of the GSTools randomisation-method
 * summators.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product path
 * (gstools_b200/) never links, imports or calls it.
 *
 * What it restates.  The arithmetic of `summate` / `summate_incompr` lives in
 * the external packages gstools-cython (>=1,<2, /root/reference/pyproject.toml:44)
 * and gstools_core (>=1.0.0, pyproject.toml:67), neither of which is vendored
 * under /root/reference.  Their published algorithm is the triple loop below;
 * the mathematical definition is the reference's own docstring:
 *   scalar : src/gstools/field/generator.py:193-206
 *            u(x) = sum_i z1_i cos(k_i.x) + z2_i sin(k_i.x)
 *   vector : src/gstools/field/generator.py:479-495
 *            u_d(x) = sum_i p_d(k_i) (z1_i cos(k_i.x) + z2_i sin(k_i.x)),
 *            p_d(k) = delta_{d0} - k_d k_0 / |k|^2
 * Argument shapes follow the call sites src/gstools/field/generator.py:42-64
 * (cov_samples (d,N), z_1 (N,), z_2 (N,), pos (d,n)); the sqrt(var/N) scale and
 * the nugget are applied by the caller (generator.py:269-270, 561-567), not here.
 *
 * Parity pin.  tests/test_oracle_golden.py checks this file against the golden
 * values the reference's own tests assert (tests/test_randmeth.py:33-71,
 * tests/test_incomprrandmeth.py:34-59, tests/test_srf.py:259-275), using mode
 * arrays produced by the reference's RandMeth in the build container
 * (tests/golden/make_golden.py).
 *
 * Loop order: points outer (parallel, one owner per point), modes inner in
 * ascending order, phase accumulated left to right over dimensions, libm cos and
 * sin called separately, plain fp64 accumulation.  Deterministic for a given libm.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_MAX_DIM 16

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

int oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* summate: out[i] = sum_j z1[j] cos(phi_ij) + z2[j] sin(phi_ij) */
int oracle_summate(const double *cov_samples, /* (dim, n_modes) row-major */
                   const double *z1, const double *z2, /* (n_modes,) */
                   const double *pos, /* (dim, n_pts) row-major */
                   int dim, int64_t n_modes, int64_t n_pts,
                   double *out, /* (n_pts,) */
                   int num_threads)
{
    if (dim < 1 || n_modes < 0 || n_pts < 0) return 1;
    int nt = pick_threads(num_threads);
    (void)nt;
#pragma omp parallel for schedule(static) num_threads(nt)
    for (int64_t i = 0; i < n_pts; ++i) {
        double acc = 0.0;
        for (int64_t j = 0; j < n_modes; ++j) {
            double phase = 0.0;
            for (int t = 0; t < dim; ++t)
                phase += cov_samples[(int64_t)t * n_modes + j] * pos[(int64_t)t * n_pts + i];
            acc += z1[j] * cos(phase) + z2[j] * sin(phase);
        }
        out[i] = acc;
    }
    return 0;
}

/* summate_incompr: out[t,i] = sum_j p_t(k_j) (z1[j] cos(phi_ij) + z2[j] sin(phi_ij)) */
int oracle_summate_incompr(const double *cov_samples, const double *z1, const double *z2,
                           const double *pos, int dim, int64_t n_modes, int64_t n_pts,
                           double *out, /* (dim, n_pts) row-major */
                           int num_threads)
{
    if (dim < 1 || dim > ORACLE_MAX_DIM || n_modes < 0 || n_pts < 0) return 1;
    int nt = pick_threads(num_threads);
    (void)nt;
#pragma omp parallel for schedule(static) num_threads(nt)
    for (int64_t i = 0; i < n_pts; ++i) {
        double acc[ORACLE_MAX_DIM];
        for (int t = 0; t < dim; ++t) acc[t] = 0.0;
        for (int64_t j = 0; j < n_modes; ++j) {
            double k2 = 0.0, phase = 0.0;
            for (int t = 0; t < dim; ++t) {
                double k = cov_samples[(int64_t)t * n_modes + j];
                k2 += k * k;
                phase += k * pos[(int64_t)t * n_pts + i];
            }
            double a = z1[j] * cos(phase) + z2[j] * sin(phase);
            double k0 = cov_samples[j];
            for (int t = 0; t < dim; ++t) {
                double e = (t == 0) ? 1.0 : 0.0;
                double proj = e - cov_samples[(int64_t)t * n_modes + j] * k0 / k2;
                acc[t] += proj * a;
            }
        }
        for (int t = 0; t < dim; ++t) out[(int64_t)t * n_pts + i] = acc[t];
    }
    return 0;
}

/* summate_fourier (next row f3): out[i] = sum_j sf[j] (z1[j] cos(phi_ij) + z2[j] sin(phi_ij)),
 * phi_ij = modes[:,j] . pos[:,i]; call site src/gstools/field/generator.py:67-75, 685-692. */
int oracle_summate_fourier(const double *spectrum_factor, const double *modes, const double *z1,
                           const double *z2, const double *pos, int dim, int64_t n_modes,
                           int64_t n_pts, double *out, int num_threads)
{
    if (dim < 1 || n_modes < 0 || n_pts < 0) return 1;
    int nt = pick_threads(num_threads);
    (void)nt;
#pragma omp parallel for schedule(static) num_threads(nt)
    for (int64_t i = 0; i < n_pts; ++i) {
        double acc = 0.0;
        for (int64_t j = 0; j < n_modes; ++j) {
            double phase = 0.0;
            for (int t = 0; t < dim; ++t)
                phase += modes[(int64_t)t * n_modes + j] * pos[(int64_t)t * n_pts + i];
            acc += spectrum_factor[j] * (z1[j] * cos(phase) + z2[j] * sin(phase));
        }
        out[i] = acc;
    }
    return 0;
}
