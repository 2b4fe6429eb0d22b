// gsb_sampler.cuh -- host-side mode-radius sampler.  No device code.
//
// RandMeth draws the radii of its wave vectors from the radial spectral density of the model
// (reference: src/gstools/field/generator.py:369-387).  Models without a closed-form inverse CDF
// (every 3-D Exponential, every Matern) go through RNG.sample_ln_pdf (src/gstools/random/rng.py:38-104):
// an emcee EnsembleSampler with 50 walkers, 20 burn-in steps and size/50*10 production steps of the
// Goodman & Weare stretch move, ~440 tiny vectorised log-pdf evaluations driven from Python -- about
// 80 ms per seed, which is what an ensemble of conditioned fields waits for once summation and kriging
// run on the GPU.  This file restates that sampler natively and STREAM-COMPATIBLY: it consumes the
// numpy legacy MT19937 stream exactly as emcee does (RandomState.choice / shuffle / rand / randint /
// rand), so for a given seed it walks the same chain and returns the same radii.
//
// Restated from the published algorithm (Goodman & Weare 2010; Foreman-Mackey et al. 2013, "emcee",
// red/blue split of the walkers, a = 2) and from numpy's documented legacy generators:
//   random_sample : (a >> 5, b >> 6) -> (a * 2^26 + b) / 2^53 from two 32-bit outputs
//   shuffle       : Fisher-Yates from the top, j = random_interval(i) by masked rejection on 32 bits
//   randint(n)    : masked rejection on 32 bits
//   choice(p=[1]) : one random_sample
// The log-pdf is the reference's `CovModel.ln_spectral_rad_pdf` (covmodel/base.py:553-560,
// covmodel/tools.py:374-406) with a native closed form for Exponential (covmodel/models.py:217-224), Matern
// (models.py:434-449) and Gaussian (models.py:147-151; 3-D has no inverse CDF); every other model's own Python
// log-pdf is called back per half ensemble (gsb_sample_radii_mcmc_cb).
#pragma once

#include <cmath>
#include <cstdint>
#include <vector>

#include "gsb_common.cuh"

namespace gsb {

struct Mt19937 {
    uint32_t key[624];
    int pos;
    void regen()
    {
        constexpr uint32_t UPPER = 0x80000000u, LOWER = 0x7fffffffu, MATRIX_A = 0x9908b0dfu;
        int kk = 0;
        uint32_t y;
        for (; kk < 624 - 397; ++kk) {
            y = (key[kk] & UPPER) | (key[kk + 1] & LOWER);
            key[kk] = key[kk + 397] ^ (y >> 1) ^ ((y & 1u) ? MATRIX_A : 0u);
        }
        for (; kk < 623; ++kk) {
            y = (key[kk] & UPPER) | (key[kk + 1] & LOWER);
            key[kk] = key[kk + (397 - 624)] ^ (y >> 1) ^ ((y & 1u) ? MATRIX_A : 0u);
        }
        y = (key[623] & UPPER) | (key[0] & LOWER);
        key[623] = key[396] ^ (y >> 1) ^ ((y & 1u) ? MATRIX_A : 0u);
        pos = 0;
    }
    uint32_t next32()
    {
        if (pos >= 624) regen();
        uint32_t y = key[pos++];
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= y >> 18;
        return y;
    }
    double next_double()
    {
        const int32_t a = (int32_t)(next32() >> 5), b = (int32_t)(next32() >> 6);
        return (a * 67108864.0 + b) / 9007199254740992.0;
    }
    // numpy.random.RandomState(seed) for an integer seed: init_genrand (legacy seeding), position at the end
    void seed(uint32_t s)
    {
        for (int i = 0; i < 624; ++i) {
            key[i] = s;
            s = 1812433253u * (s ^ (s >> 30)) + (uint32_t)i + 1u;
        }
        pos = 624;
    }
    // uniform integer in [0, max] by masked rejection (numpy random_interval / bounded_masked_uint32)
    uint32_t bounded(uint32_t max)
    {
        if (max == 0) return 0;
        uint32_t mask = max;
        mask |= mask >> 1;
        mask |= mask >> 2;
        mask |= mask >> 4;
        mask |= mask >> 8;
        mask |= mask >> 16;
        uint32_t v;
        while ((v = (next32() & mask)) > max) {
        }
        return v;
    }
};

// numpy's `array ** scalar` fast paths, then libm pow
static inline double np_pow(double x, double e)
{
    if (e == 2.0) return x * x;
    if (e == 1.0) return x;
    if (e == 0.5) return std::sqrt(x);
    if (e == -1.0) return 1.0 / x;
    if (e == 0.0) return 1.0;
    return std::pow(x, e);
}

struct RadPdf {
    int kind;   // GSB_PDF_EXPONENTIAL / GSB_PDF_MATERN / GSB_PDF_GAUSSIAN
    int dim;
    double len_rescaled, nu;
    // factors of the densities that do not depend on k: the SAME sub-expressions the per-call formulas evaluate,
    // computed once (bit-identical results; pow / tgamma / lgamma were 80 % of a chain's time)
    double c_fac = 0.0, c_exp = 0.0, c_lg1 = 0.0, c_lg2 = 0.0, c_lg3 = 0.0, c_rad = 0.0;

    RadPdf(int kind_, int dim_, double len_, double nu_) : kind(kind_), dim(dim_), len_rescaled(len_), nu(nu_)
    {
        const double l = len_rescaled;
        if (kind == GSB_PDF_GAUSSIAN) {
            c_fac = std::pow(l / 2.0 / std::sqrt(M_PI), dim);
        } else if (kind == GSB_PDF_EXPONENTIAL) {
            c_fac = std::pow(l, dim) * std::tgamma((dim + 1) / 2.0);
            c_exp = (dim + 1) / 2.0;
        } else {
            c_fac = std::pow(l / std::sqrt(M_PI), dim);
            c_exp = -(nu + dim / 2.0);
            c_lg1 = std::lgamma(nu + dim / 2.0);
            c_lg2 = std::lgamma(nu);
            c_lg3 = dim * std::log(std::sqrt(nu));
        }
        if (dim > 3) c_rad = std::pow(std::sqrt(M_PI), dim) / std::tgamma(dim / 2.0 + 1);
    }

    double rad_fac(double r) const   // covmodel/tools.py:118-140
    {
        if (dim == 1) return 2.0;
        if (dim == 2) return (2 * M_PI) * r;
        if (dim == 3) return (4 * M_PI) * (r * r);
        return dim * std::pow(r, dim - 1) * std::pow(std::sqrt(M_PI), dim) / std::tgamma(dim / 2.0 + 1);
    }
    double spectral_density(double k) const
    {
        const double l = len_rescaled;
        if (kind == GSB_PDF_GAUSSIAN) {    // models.py:147-151
            const double h = k * l / 2.0;
            return c_fac * std::exp(-(h * h));
        }
        if (kind == GSB_PDF_EXPONENTIAL)   // models.py:217-224
            return c_fac / np_pow(M_PI * (1.0 + (k * l) * (k * l)), c_exp);
        const double x = (k * l) * (k * l);   // Matern, models.py:434-449
        if (nu > 20.0)
            return c_fac * std::exp(-x) * (1 + 0.5 * (x * x) / nu) * std::pow(std::sqrt(1 + x / nu), -dim);
        return c_fac * std::exp(c_exp * std::log(1.0 + x / nu) + c_lg1 - c_lg2 - c_lg3);
    }
    // np.log(spectral_rad_pdf(model, r)), covmodel/tools.py:374-406 and covmodel/base.py:557-560
    double ln_pdf(double r) const
    {
        r = std::fabs(r);
        double res;
        if (dim > 1 && r <= 1e-8) res = 0.0;                       // np.isclose(r, 0)
        else res = rad_fac(r) * std::fabs(spectral_density(r));
        if (!std::isfinite(res)) res = 0.0;
        res = std::fmax(res, 0.0);
        return std::log(res);                                       // log(0) = -inf, as np.log under errstate
    }
};

// One EnsembleSampler.run_mcmc(state, nsteps) with the default StretchMove; coords / logp are updated in
// place, `chain` (nsteps x nwalkers) receives the positions after every step when non-null.
// Returns 0, 1 when a proposal or its log-pdf is not finite (emcee raises ValueError there), 2 when `eval` aborted.
//
// `eval(q, n, out)` evaluates the log-pdf of n proposals at once (emcee runs the reference's sampler with
// vectorize=True: one call per half ensemble); it returns non-zero to abort.
template <typename Eval>
static inline int stretch_run(Eval &&eval, Mt19937 &rng, int nwalkers, int nsteps, double *coords, double *logp,
                              double *chain)
{
    std::vector<int> inds(nwalkers), S, C;
    std::vector<double> zz, q, nlp;
    std::vector<unsigned char> accepted(nwalkers);
    for (int step = 0; step < nsteps; ++step) {
        (void)rng.next_double();                                    // move = random.choice(moves, p=weights)
        for (int i = 0; i < nwalkers; ++i) inds[i] = i % 2;
        for (int i = nwalkers - 1; i >= 1; --i) {                   // random.shuffle(inds)
            const uint32_t j = rng.bounded((uint32_t)i);
            const int tmp = inds[i];
            inds[i] = inds[j];
            inds[j] = tmp;
        }
        std::fill(accepted.begin(), accepted.end(), 0);
        for (int split = 0; split < 2; ++split) {
            S.clear();
            C.clear();
            for (int i = 0; i < nwalkers; ++i) (inds[i] == split ? S : C).push_back(i);
            const int ns = (int)S.size(), nc = (int)C.size();
            zz.resize(ns);
            q.resize(ns);
            nlp.resize(ns);
            for (int k = 0; k < ns; ++k) {                          // zz = ((a - 1) * rand(ns) + 1) ** 2 / a
                const double t = 1.0 * rng.next_double() + 1;
                zz[k] = (t * t) / 2.0;
            }
            for (int k = 0; k < ns; ++k) {                          // rint = randint(nc, size=ns)
                const int r = (int)rng.bounded((uint32_t)(nc - 1));
                const double c = coords[C[r]];
                q[k] = c - (c - coords[S[k]]) * zz[k];
            }
            for (int k = 0; k < ns; ++k)
                if (!std::isfinite(q[k])) return 1;
            if (ns > 0 && eval(q.data(), ns, nlp.data())) return 2;
            for (int k = 0; k < ns; ++k)
                if (std::isnan(nlp[k])) return 1;
            for (int k = 0; k < ns; ++k) {                          // ndim = 1: factors = 0 * log(zz)
                const double lnpdiff = 0.0 * std::log(zz[k]) + nlp[k] - logp[S[k]];
                if (lnpdiff > std::log(rng.next_double())) accepted[S[k]] = 1;
            }
            for (int k = 0; k < ns; ++k)
                if (accepted[S[k]]) {
                    coords[S[k]] = q[k];
                    logp[S[k]] = nlp[k];
                }
        }
        if (chain)
            for (int i = 0; i < nwalkers; ++i) chain[(size_t)step * nwalkers + i] = coords[i];
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Mode sets of MANY seeds at once (ensembles): the random streams of RandMeth.reset_seed
// (src/gstools/field/generator.py:346-387) restated natively, one seed per task, tasks spread over host threads.
//   RNG(seed)             -> MasterRNG(seed): RandomState(seed).randint(1, 2**16) per access of RNG.random
//                            (random/tools.py:30-35, random/rng.py:193-203)
//   z_1, z_2              -> RandomState(s).normal(size=N): numpy's legacy polar Box-Muller with its cached second value
//   sample_sphere         -> uniform(0, 2 pi, N) [and uniform(-1, 1, N) in 3-D] (rng.py:163-174); the trigonometry
//                            stays in numpy (its SIMD cos / sin are not libm's) and runs on the whole batch at once
//   sample_ln_pdf         -> rand(nwalkers), two generator states, the stretch-move chain above, choice(chain, N)
// Outputs per seed: z1[N], z2[N], ang1[N], ang2[N] (3-D only), rad[N].
// ---------------------------------------------------------------------------------------------
struct LegacyGauss {     // numpy legacy_gauss
    Mt19937 &rng;
    bool has = false;
    double stored = 0.0;
    explicit LegacyGauss(Mt19937 &r) : rng(r) {}
    double next()
    {
        if (has) {
            has = false;
            return stored;
        }
        double x1, x2, r2;
        do {
            x1 = 2.0 * rng.next_double() - 1.0;
            x2 = 2.0 * rng.next_double() - 1.0;
            r2 = x1 * x1 + x2 * x2;
        } while (r2 >= 1.0 || r2 == 0.0);
        const double f = std::sqrt(-2.0 * std::log(r2) / r2);
        stored = f * x1;
        has = true;
        return f * x2;
    }
};

struct ModeBatchArgs {
    RadPdf pdf;
    int dim;
    int64_t mode_no;
    int nwalkers, burn_in;
    int64_t n_steps;
    double sample_around, two_pi;
};

// one seed; returns 0, or 1 when the chain hit a non-finite proposal / log-pdf
static inline int sample_modes_one(const ModeBatchArgs &a, uint32_t seed, double *z1, double *z2, double *ang1,
                                   double *ang2, double *rad, std::vector<double> &chain)
{
    Mt19937 master, rs;
    master.seed(seed);
    auto next_stream = [&]() { rs.seed(1u + master.bounded(65534u)); };     // RandomState(randint(1, 2**16))
    const int64_t N = a.mode_no;
    next_stream();
    {
        LegacyGauss g(rs);
        for (int64_t k = 0; k < N; ++k) z1[k] = g.next();
    }
    next_stream();
    {
        LegacyGauss g(rs);
        for (int64_t k = 0; k < N; ++k) z2[k] = g.next();
    }
    if (a.dim == 1) {            // choice([-1, 1], size): randint(0, 2)
        next_stream();
        for (int64_t k = 0; k < N; ++k) ang1[k] = rs.bounded(1u) ? 1.0 : -1.0;
    } else {
        next_stream();
        for (int64_t k = 0; k < N; ++k) ang1[k] = 0.0 + a.two_pi * rs.next_double();
        if (a.dim == 3) {
            next_stream();
            for (int64_t k = 0; k < N; ++k) ang2[k] = -1.0 + 2.0 * rs.next_double();
        }
    }
    // RNG.sample_ln_pdf (rng.py:72-104)
    std::vector<double> coords((size_t)a.nwalkers), logp((size_t)a.nwalkers);
    next_stream();
    for (int i = 0; i < a.nwalkers; ++i) coords[(size_t)i] = rs.next_double() * a.sample_around;
    auto eval = [&](const double *q, int n, double *out) -> int {
        for (int k = 0; k < n; ++k) out[k] = a.pdf.ln_pdf(q[k]);
        return 0;
    };
    eval(coords.data(), a.nwalkers, logp.data());
    for (int i = 0; i < a.nwalkers; ++i)
        if (!std::isfinite(coords[(size_t)i]) || std::isnan(logp[(size_t)i])) return 1;
    chain.resize((size_t)a.n_steps * a.nwalkers);
    next_stream();
    if (stretch_run(eval, rs, a.nwalkers, a.burn_in, coords.data(), logp.data(), nullptr)) return 1;
    next_stream();
    if (stretch_run(eval, rs, a.nwalkers, (int)a.n_steps, coords.data(), logp.data(), chain.data())) return 1;
    next_stream();
    const uint32_t pop = (uint32_t)chain.size();
    for (int64_t k = 0; k < N; ++k) rad[k] = chain[rs.bounded(pop - 1u)];       // choice(samples, size): randint(0, pop)
    return 0;
}

}  // namespace gsb
