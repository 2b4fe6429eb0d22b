// gsb_direct.cuh -- direct (unstructured) summation kernels for sm_100a.
//
// Replaces the native loop nest behind gstools.field.generator._summate /
// _summate_incompr (reference: src/gstools/field/generator.py:42-64; math at
// generator.py:193-199 and 479-495) for a flat (dim, n) position array.
//
// Design (FP64-pipe bound; see DESIGN.md "direct kernel"):
//   * one thread owns P points (positions in registers) and accumulates them over ALL modes in
//     ascending mode order: no atomics, no cross-thread reduction, bit-reproducible;
//   * the CTA streams tiles of packed mode records through shared memory with
//     cp.async.bulk + mbarrier (TMA unit), double buffered; every lane of a warp reads the
//     same record, so the LDS are broadcasts;
//   * sin/cos: the wave vectors are pre-scaled to SIXTEENTH turns (k * 8/pi), so that
//     u = k.x, n = rint(u) (magic-number add), r = u - n in [-1/2, 1/2] is an EXACT
//     reduction with 3 DADD; sin(pi/8 r) and cos(pi/8 r) are degree-7 / degree-6 polynomials
//     (sincos_coeffs.cuh, abs error 9.3e-15 / 4.3e-13; tools/gen_sincos_coeffs.py lists the
//     accuracy / table-size trade-off: every halving of the reduced range saves one DFMA per
//     polynomial and doubles the table below);
//   * the rotation by n sixteenth turns costs no FP64 and no select: each record carries the 16
//     rotated weight pairs W_q = (z1 c_q + z2 s_q, z2 c_q - z1 s_q), c_q + i s_q = exp(i q pi/8), as two
//     arrays of 16 doubles -- each exactly one 128-byte row of shared memory, so the per-lane LDS.64
//     with index (n & 15) are conflict free;
//   * FP64 instructions per (point, mode): D + 13 (scalar), D + 13 + D (incompressible).
#pragma once

#include "gsb_common.cuh"
#include "sincos_coeffs.cuh"

namespace gsb {

#ifndef GSB_DIRECT_TM
#define GSB_DIRECT_TM 64
#endif
constexpr int DIRECT_TM = GSB_DIRECT_TM;
#ifndef GSB_DIRECT_UNROLL
#define GSB_DIRECT_UNROLL 2
#endif
constexpr int DIRECT_UNROLL = GSB_DIRECT_UNROLL;   // modes per unrolled iteration of the inner loop   // modes per shared-memory tile (2 buffers x 64 x 40 doubles <= 48 KB static)
constexpr double RINT_MAGIC = 6755399441055744.0;  // 1.5 * 2^52
constexpr double EIGHT_OVER_PI = 2.54647908947032537230;
constexpr int DIRECT_NROT = 16;              // rotations per record (sixteenth turns)
constexpr int DIRECT_NW = 2 * DIRECT_NROT;   // doubles of rotated weights per record: Wx[16], Wy[16]

__host__ __device__ constexpr int direct_koff(int D) { return (D + 1) & ~1; }
__host__ __device__ constexpr int direct_rec(int D, bool vec)
{
    return direct_koff(D) + DIRECT_NW + (vec ? direct_koff(D) : 0);
}

// ---------------------------------------------------------------------------------------------
// mode packing: (cov_samples, z1, z2) -> records [kq_0..kq_{D-1} pad | Wx[16] | Wy[16] | p_0..p_{D-1} pad]
// kq = k * 8/pi (sixteenth turns); (Wx[q], Wy[q]) = weight pair after rotating by q sixteenth turns;
// p_t = delta_t0 - k_t k_0 / |k|^2 (incompressible projector, generator.py:479-495).
// Records j >= n_modes (padding up to n_modes_pad) are all-zero and contribute exactly 0.
// ---------------------------------------------------------------------------------------------
__global__ void pack_modes_kernel(const double *__restrict__ cov, const double *__restrict__ z1,
                                  const double *__restrict__ z2, const double *__restrict__ sf,
                                  int dim, int64_t n_modes, int64_t n_modes_pad, int vec,
                                  double *__restrict__ recs)
{
    const int koff = direct_koff(dim);
    const int rec = direct_rec(dim, vec != 0);
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n_modes_pad;
         j += (int64_t)gridDim.x * blockDim.x) {
        double *R = recs + j * rec;
        if (j >= n_modes) {
            for (int t = 0; t < rec; ++t) R[t] = 0.0;
            continue;
        }
        double k2 = 0.0;
        for (int t = 0; t < dim; ++t) {
            double k = cov[(int64_t)t * n_modes + j];
            R[t] = k * EIGHT_OVER_PI;
            k2 += k * k;
        }
        for (int t = dim; t < koff; ++t) R[t] = 0.0;
        // sf: optional per-mode spectrum factor of the Fourier generator (generator.py:685-692)
        const double w = sf ? sf[j] : 1.0;
        const double a = w * z1[j], b = w * z2[j];
        double *W = R + koff;
        // z1 cos(phi) + z2 sin(phi) with phi = (q + r) pi/8:  cos(r pi/8) * Wx[q] + sin(r pi/8) * Wy[q]
        for (int q = 0; q < DIRECT_NROT; ++q) {
            const double cq = cospi(q / 8.0), sq = sinpi(q / 8.0);   // exact 0, +-1 at the quarter turns
            W[q] = a * cq + b * sq;
            W[DIRECT_NROT + q] = b * cq - a * sq;
        }
        if (vec) {
            double *Pj = W + DIRECT_NW;
            const double k0 = cov[j];
            for (int t = 0; t < dim; ++t) {
                const double e = (t == 0) ? 1.0 : 0.0;
                Pj[t] = e - cov[(int64_t)t * n_modes + j] * k0 / k2;
            }
            for (int t = dim; t < koff; ++t) Pj[t] = 0.0;
        }
    }
}

struct DirectParams {
    const double *recs;     // packed mode records, n_modes_pad * REC doubles
    int64_t n_modes_pad;    // multiple of 4
    const double *pos;      // (D, n_pts), row stride pos_ld
    int64_t pos_ld;
    int64_t n_pts;
    double *out;            // scalar: (n_pts,)   vector: (D, n_pts) row stride out_ld
    int64_t out_ld;
    int n_split;            // mode splits (gridDim.y); >1 writes the per-tile sums to `partial`
    double *partial;        // (n_tiles, ncomp, n_pts) when n_split > 1
    Epi epi;                // fused caller epilogue (off: raw sums)
};

// residual sin/cos on r in [-1/2, 1/2] sixteenth turns: returns ps = sin(pi/8 r)/r and pc = cos(pi/8 r)
__device__ __forceinline__ void qt_polys(double z, double &ps, double &pc)
{
    double s = fma(z, GSB_S3, GSB_S2);
    double c = fma(z, GSB_C3, GSB_C2);
    s = fma(z, s, GSB_S1);
    c = fma(z, c, GSB_C1);
    ps = fma(z, s, GSB_S0);
    pc = fma(z, c, GSB_C0);
}

template <int D, int P, bool VEC, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) direct_kernel(const DirectParams prm)
{
    constexpr int KOFF = direct_koff(D);
    constexpr int REC = direct_rec(D, VEC);
    constexpr int NC = VEC ? D : 1;
    constexpr int TM = DIRECT_TM;

    __shared__ __align__(128) double tile[2][TM * REC];
    __shared__ __align__(8) uint64_t full[2];

    const int tid = threadIdx.x;
    const int64_t base = (int64_t)blockIdx.x * (THREADS * P) + tid;

    // Summation order of a point (the same in every launch configuration AND with the mode loop split over CTAs):
    // modes are summed in ascending order inside a tile of 64 (`tacc`), the tile sums are added in ascending order
    // (`acc`).  A CTA that owns only some tiles stores their sums and reduce_partials_kernel adds them in the same
    // order, so a point's bits depend neither on what else is in the call nor on how the work was cut.
    double x[P][D];
    double acc[P][NC], tacc[P][NC];
#pragma unroll
    for (int p = 0; p < P; ++p) {
        const int64_t i = base + (int64_t)p * THREADS;
#pragma unroll
        for (int t = 0; t < D; ++t) x[p][t] = (i < prm.n_pts) ? prm.pos[t * prm.pos_ld + i] : 0.0;
#pragma unroll
        for (int c = 0; c < NC; ++c) acc[p][c] = 0.0;
    }

    // contiguous range of mode tiles handled by this CTA (all of them unless n_split > 1)
    const int n_tiles = (int)((prm.n_modes_pad + TM - 1) / TM);
    const int per = (n_tiles + prm.n_split - 1) / prm.n_split;
    const int t0 = blockIdx.y * per;
    const int t1 = min(n_tiles, t0 + per);

    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        fence_barrier_init();
    }
    __syncthreads();

    auto issue = [&](int t, int b) {
        const int64_t j0 = (int64_t)t * TM;
        const int cnt = (int)min((int64_t)TM, prm.n_modes_pad - j0);
        const uint32_t bytes = (uint32_t)(cnt * REC * sizeof(double));
        mbar_arrive_expect_tx(&full[b], bytes);
        bulk_g2s(&tile[b][0], prm.recs + j0 * REC, bytes, &full[b]);
    };
    if (tid == 0 && t0 < t1) issue(t0, 0);

    for (int t = t0; t < t1; ++t) {
        const int b = (t - t0) & 1;
        if (tid == 0 && t + 1 < t1) issue(t + 1, b ^ 1);
        mbar_wait(&full[b], ((t - t0) >> 1) & 1);

        const int cnt = (int)min((int64_t)TM, prm.n_modes_pad - (int64_t)t * TM);
        const char *T = reinterpret_cast<const char *>(&tile[b][0]);
#pragma unroll
        for (int p = 0; p < P; ++p)
#pragma unroll
            for (int c = 0; c < NC; ++c) tacc[p][c] = 0.0;
#pragma unroll DIRECT_UNROLL
        for (int j = 0; j < cnt; ++j) {
            const double *R = reinterpret_cast<const double *>(T + j * (REC * 8));
            double k[D];
#pragma unroll
            for (int t2 = 0; t2 < D; ++t2) k[t2] = R[t2];
            double pj[NC];
            if (VEC) {
#pragma unroll
                for (int c = 0; c < NC; ++c) pj[c] = R[KOFF + DIRECT_NW + c];
            }
#pragma unroll
            for (int p = 0; p < P; ++p) {
                double u = k[0] * x[p][0];
#pragma unroll
                for (int t2 = 1; t2 < D; ++t2) u = fma(k[t2], x[p][t2], u);
                // (rint on the conversion pipe -- F2I + I2F instead of two of the DADDs -- measured +2 % only:
                // 64-bit conversions are not cheaper than DADDs here, and they narrow the phase domain)
                const double v = u + RINT_MAGIC;           // rint(u) lands in the low mantissa bits
                const int q = lo32(v);
                const double r = u - (v - RINT_MAGIC);     // exact, |r| <= 1/2
                double2 w;
                w.x = R[KOFF + (q & (DIRECT_NROT - 1))];
                w.y = R[KOFF + DIRECT_NROT + (q & (DIRECT_NROT - 1))];
                const double z = r * r;
                double ps, pc;
                qt_polys(z, ps, pc);
                const double wr = w.y * r;
                if (!VEC) {
                    tacc[p][0] = fma(wr, ps, tacc[p][0]);
                    tacc[p][0] = fma(w.x, pc, tacc[p][0]);
                } else {
                    const double a = fma(w.x, pc, wr * ps);
#pragma unroll
                    for (int c = 0; c < NC; ++c) tacc[p][c] = fma(pj[c], a, tacc[p][c]);
                }
            }
        }
        if (prm.n_split == 1) {
#pragma unroll
            for (int p = 0; p < P; ++p)
#pragma unroll
                for (int c = 0; c < NC; ++c) acc[p][c] = __dadd_rn(acc[p][c], tacc[p][c]);
        } else {
#pragma unroll
            for (int p = 0; p < P; ++p) {
                const int64_t i = base + (int64_t)p * THREADS;
                if (i < prm.n_pts) {
#pragma unroll
                    for (int c = 0; c < NC; ++c) prm.partial[((int64_t)t * NC + c) * prm.n_pts + i] = tacc[p][c];
                }
            }
        }
        __syncthreads();  // everyone is done with buffer b before it is refilled
    }
    if (prm.n_split > 1) return;

#pragma unroll
    for (int p = 0; p < P; ++p) {
        const int64_t i = base + (int64_t)p * THREADS;
        if (i < prm.n_pts) {
#pragma unroll
            for (int c = 0; c < NC; ++c) prm.out[c * prm.out_ld + i] = epi_apply(prm.epi, acc[p][c], c, i);
        }
    }
}

// the tile sums of a mode-split launch added in ascending tile order: the unsplit kernel's own order and bits
__global__ void reduce_partials_kernel(const double *__restrict__ partial, int n_tiles, int ncomp,
                                       int64_t n_pts, double *__restrict__ out, int64_t out_ld,
                                       const Epi epi)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int c = blockIdx.y;
    if (i >= n_pts) return;
    double s = 0.0;
    for (int k = 0; k < n_tiles; ++k) s = __dadd_rn(s, partial[((int64_t)k * ncomp + c) * n_pts + i]);
    out[c * out_ld + i] = epi_apply(epi, s, c, i);
}

// ---------------------------------------------------------------------------------------------
// launcher
// ---------------------------------------------------------------------------------------------
template <int D, int P, bool VEC, int THREADS, int MINB>
inline void launch_direct_cfg(const DirectParams &prm, cudaStream_t st)
{
    dim3 grid((unsigned)((prm.n_pts + (int64_t)THREADS * P - 1) / ((int64_t)THREADS * P)),
              (unsigned)prm.n_split);
    direct_kernel<D, P, VEC, THREADS, MINB><<<grid, THREADS, 0, st>>>(prm);
    g_launches.fetch_add(1);
}

template <int D, bool VEC>
inline void launch_direct_dim(const DirectParams &prm, int cfg, cudaStream_t st)
{
    if constexpr (D <= 4) {
#ifndef GSB_DIRECT_MINB
#define GSB_DIRECT_MINB 2
#endif
        // 8 points per thread, 8 warps per SM: fewer warps with more independent chains keep the FP64
        // pipe fuller than 4 points x 16 warps (+3 % 2-D, +4 % 3-D, +8 % incompressible; tools/direct_cfg_sweep.py)
        if constexpr (D <= 3) {
            if (cfg >= 2) return launch_direct_cfg<D, 8, VEC, 128, 2>(prm, st);
        }
        if (cfg >= 2) return launch_direct_cfg<D, 4, VEC, 256, GSB_DIRECT_MINB>(prm, st);
    }
    if (cfg >= 1) launch_direct_cfg<D, 2, VEC, 128, 4>(prm, st);
    else launch_direct_cfg<D, 1, VEC, 64, 8>(prm, st);
}

// points per CTA of configuration cfg
inline int64_t direct_cfg_points(int cfg, int dim)
{
    if (cfg >= 2 && dim > 4) cfg = 1;
    return cfg >= 2 ? 1024 : (cfg == 1 ? 256 : 64);
}

inline int launch_direct(int dim, bool vec, const DirectParams &prm, int cfg, cudaStream_t st)
{
#define GSB_DIRECT_CASE(DD)                                                                    \
    case DD:                                                                                   \
        if (vec) {                                                                             \
            if constexpr (DD == 2 || DD == 3) launch_direct_dim<DD, true>(prm, cfg, st);       \
            else return fail(GSB_ERR_ARGUMENT, "vector field needs dim 2 or 3");               \
        } else {                                                                               \
            launch_direct_dim<DD, false>(prm, cfg, st);                                        \
        }                                                                                      \
        break;
    switch (dim) {
        GSB_DIRECT_CASE(1)
        GSB_DIRECT_CASE(2)
        GSB_DIRECT_CASE(3)
        GSB_DIRECT_CASE(4)
        GSB_DIRECT_CASE(5)
        GSB_DIRECT_CASE(6)
        GSB_DIRECT_CASE(7)
        GSB_DIRECT_CASE(8)
    default:
        return fail(GSB_ERR_ARGUMENT, "dim must be in 1..8");
    }
#undef GSB_DIRECT_CASE
    GSB_CUDA(cudaGetLastError());
    return GSB_OK;
}

}  // namespace gsb
