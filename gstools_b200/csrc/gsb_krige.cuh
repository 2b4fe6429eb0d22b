// gsb_krige.cuh -- kriging evaluation on sm_100a.
//
// Replaces the native `calc_field_krige_and_variance` / `calc_field_krige` of gstools-cython /
// gstools_core (reference: imported src/gstools/krige/base.py:16-19, 30-33; dispatched :42-61; called
// from Krige._summate :307-317):
//     field[k] = sum_i cond[i] (M kv)[i,k]        error[k] = sum_i kv[i,k] (M kv)[i,k]
// with M = krig_mat (K,K), kv = krig_vecs (K,n).  The reference evaluates the full product M kv
// (K^2 FMAs per point).  Here both results come from exact algebraic regroupings:
//     field[k] = sum_j w[j] kv[j,k],                 w = M^T cond                       (K FMAs per point)
//     error[k] = sum_j kv[j,k] sum_{i<=j} L[j,i] kv[i,k],   L[j,j] = M[j,j], L[j,i] = M[i,j] + M[j,i]
// i.e. the quadratic form through a LOWER-TRIANGULAR operand: about K^2/2 FMAs per point.  The row w
// rides along as row K of the same operand (a padding row of the last 128-row tile), so the field
// falls out of the same contraction.  The identities hold for any M (no symmetry assumed); only the
// rounding differs from the reference's order (tests state the bound).
//
// Kernels:
//   krige_w_kernel       w = M^T cond
//   krige_tiles_kernel   the operand [L; w] PRE-TILED in the shared-memory layout of a pipeline stage
//                        ([row tile r][depth stage s <= 8(r+1)] -> 128 x 20 doubles), O(K^2), once per call
//   krige_kernel         persistent, one CTA of 8 warps per SM.  Work unit = (column tile of 128 points,
//                        pair of row tiles {R-1-p, p}): every unit costs the same (R+1)*8 depth stages, so
//                        static striding over units is balanced.  Per stage of 16 depth indices one bulk
//                        copy of the 20 KB operand tile plus 16 bulk copies of 1 KB -- the rows of kv
//                        straight from the caller's row-major array into the padded B layout -- land in a
//                        5-slot mbarrier ring.  Consumers: the DMMA.8x8x4 loop of the separable
//                        contraction (gsb_separable.cuh).  Epilogue: the accumulators times the matching kv
//                        entries -- read from the stage buffers of the diagonal stages, where they pass through
//                        shared memory anyway -- reduced over rows in a fixed order (registers -> shuffles ->
//                        shared memory), one partial per (pair, point); row K is the field.
//   krige_finish_kernel  error[k] = sum_p partial[p][k] in ascending p (deterministic, no atomics).
//   krige_field_kernel   `calc_field_krige` (no variance): field = w . kv, one pass over kv (HBM bound).
#pragma once

#include "gsb_separable.cuh"

namespace gsb {

#ifndef GSB_KRG_STAGES
#define GSB_KRG_STAGES 5
#endif
constexpr int KRG_STAGES = GSB_KRG_STAGES;         // pipeline ring slots
constexpr int KRG_STAGE_DOUBLES = SEP_A_TILE + SEP_B_TILE;
constexpr size_t KRG_SMEM_BYTES = (size_t)KRG_STAGES * KRG_STAGE_DOUBLES * sizeof(double) + 2 * KRG_STAGES * sizeof(uint64_t) + 128;
constexpr int KRG_KD = 2 * SEP_KC;                 // depth indices per pipeline stage (16)
constexpr int KRG_SPT = SEP_TM / KRG_KD;           // depth stages per 128-row tile (8)
static_assert(KRG_KD % SEP_WARPS == 0, "kv rows of a stage are split evenly over the warps");
static_assert(SEP_TM % KRG_KD == 0, "row tile must be a whole number of depth stages");

// first operand tile of row tile r (each row tile r owns KRG_SPT*(r+1) tile slots)
__host__ __device__ inline int64_t krige_tile_off(int r) { return (int64_t)KRG_SPT * r * (r + 1) / 2; }

struct KrigeParams {
    const double *atile;    // pre-tiled operand [L; w]
    const double *kv;       // (K, n), row stride ld; 16-byte aligned base, ld and n even
    int64_t ld;
    int64_t n;
    int64_t n_copy;         // n rounded up to even: columns the bulk copies may read
    int K;                  // kriging system size
    int R;                  // 128-row tiles of the operand (covers rows 0..K, row K = w)
    int n_pairs;            // (R + 1) / 2
    int n_dstages;          // ceil(K / 16): depth stages holding real rows
    int64_t n_col_tiles;
    const double *zeros;    // >= SEP_TN doubles of 0.0 (source for depth rows >= K)
    const double *btile;    // TILED variant: kv pre-tiled by kvgen_kernel, tile (ct, s) = [16][SEP_BST] doubles at
                            // ((ct * n_dstages + s) * SEP_B_TILE); rows >= K are zero
    double *partial;        // (n_pairs, n)
    double *field;          // (n,)
};

__global__ void krige_w_kernel(const double *__restrict__ mat, const double *__restrict__ cond, int K,
                               double *__restrict__ w)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= K) return;
    double s = 0.0;
    for (int l = 0; l < K; ++l) s = fma(mat[(int64_t)l * K + i], cond[l], s);
    w[i] = s;
}

// one CTA per operand tile (r, s)
__global__ void krige_tiles_kernel(const double *__restrict__ mat, const double *__restrict__ w, int K,
                                   int R, double *__restrict__ atile)
{
    int r = 0;
    int64_t rem = blockIdx.x;
    while (rem >= (int64_t)KRG_SPT * (r + 1)) { rem -= (int64_t)KRG_SPT * (r + 1); ++r; }
    const int s = (int)rem;
    double *T = atile + (krige_tile_off(r) + s) * SEP_A_TILE;
    for (int e = threadIdx.x; e < SEP_A_TILE; e += blockDim.x) {
        const int row = e / SEP_AST, d = e % SEP_AST;
        const int j = r * SEP_TM + row;      // operand row
        const int i = s * KRG_KD + d;        // depth index
        double v = 0.0;
        if (d < KRG_KD && i < K) {
            if (j < K) {
                if (i < j) v = mat[(int64_t)i * K + j] + mat[(int64_t)j * K + i];
                else if (i == j) v = mat[(int64_t)j * K + j];
            } else if (j == K) {
                v = w[i];
            }
        }
        T[e] = v;
    }
}

// TILED = false: kv is the caller's row-major array (16 row copies per stage)
// TILED = true : kv was generated on the device, pre-tiled in the stage layout (one copy per stage)
template <bool TILED>
__global__ void __launch_bounds__(SEP_THREADS, 1) krige_kernel(const KrigeParams prm)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *stage_base = reinterpret_cast<double *>(smem_raw);
    uint64_t *full = reinterpret_cast<uint64_t *>(stage_base + KRG_STAGES * KRG_STAGE_DOUBLES);
    uint64_t *empty = full + KRG_STAGES;
    __shared__ double red[4][SEP_TN];

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    const int64_t n_units = prm.n_col_tiles * prm.n_pairs;

    if (tid == 0) {
        for (int s = 0; s < KRG_STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], SEP_WARPS);
        }
        fence_barrier_init();
    }
    __syncthreads();

    // cursor over (unit, half, stage); `half` 0 is the heavy row tile R-1-p, half 1 the light one p
    struct Cursor {
        int64_t unit;
        int half, s, r, S;
        int64_t c;
    };
    auto cur_set = [&](Cursor &q) {   // derive (c, r, S) from (unit, half)
        q.c = q.unit / prm.n_pairs;
        const int p = (int)(q.unit % prm.n_pairs);
        q.r = q.half == 0 ? prm.R - 1 - p : p;
        q.S = min(KRG_SPT * (q.r + 1), prm.n_dstages);
    };
    auto cur_advance = [&](Cursor &q, int64_t stride) {   // next stage; uniform over the CTA
        if (++q.s < q.S) return;
        q.s = 0;
        const int p = (int)(q.unit % prm.n_pairs);
        if (q.half == 0 && p != prm.R - 1 - p) {
            q.half = 1;
        } else {
            q.half = 0;
            q.unit += stride;
        }
        if (q.unit < n_units) cur_set(q);
    };

    constexpr int DEPTH = KRG_STAGES - 2;
    Cursor pf{(int64_t)blockIdx.x, 0, 0, 0, 0, 0};
    int pf_slot = 0;
    uint32_t pf_round = 0;
    if (pf.unit < n_units) cur_set(pf);
#ifndef GSB_KRG_ISSUE
#define GSB_KRG_ISSUE 1
#endif
    // Prefetch of one stage = the 20 KB operand tile + 16 kv rows of 1 KB (a bulk copy is one
    // uniform-datapath instruction per row).  A complete_tx landing before the arrive.expect_tx is fine:
    // the phase cannot complete before that one pending arrival.
    //   GSB_KRG_ISSUE 1: the warp whose turn it is issues everything, one copy per lane (17 lanes)
    //   GSB_KRG_ISSUE 2: lane 0 of every warp copies two kv rows, the turn warp adds the operand tile
    //   GSB_KRG_ISSUE 3: lane 0 of the turn warp issues all 17 copies back to back
    auto pf_issue = [&](bool lead) {
        double *A = stage_base + pf_slot * KRG_STAGE_DOUBLES;
        const int64_t col0 = pf.c * SEP_TN;
        const uint32_t row_bytes = (uint32_t)min((int64_t)SEP_TN, prm.n_copy - col0) * sizeof(double);
        const uint32_t tx = SEP_A_TILE * sizeof(double) + KRG_KD * row_bytes;
        const double *asrc = prm.atile + (krige_tile_off(pf.r) + pf.s) * SEP_A_TILE;
        if (TILED) {
            if (!lead || lane != 0) return;
            if (pf_round > 0) mbar_wait(&empty[pf_slot], (pf_round - 1) & 1);
            mbar_arrive_expect_tx(&full[pf_slot], (SEP_A_TILE + SEP_B_TILE) * sizeof(double));
            bulk_g2s(A, asrc, SEP_A_TILE * sizeof(double), &full[pf_slot]);
            bulk_g2s(A + SEP_A_TILE, prm.btile + (pf.c * prm.n_dstages + pf.s) * SEP_B_TILE,
                     SEP_B_TILE * sizeof(double), &full[pf_slot]);
            return;
        }
#if GSB_KRG_ISSUE == 1
        if (!lead) return;
        if (lane == 0) {
            if (pf_round > 0) mbar_wait(&empty[pf_slot], (pf_round - 1) & 1);
            mbar_arrive_expect_tx(&full[pf_slot], tx);
        }
        __syncwarp();
        if (lane < KRG_KD) {
            const int row = pf.s * KRG_KD + lane;
            const double *src = row < prm.K ? prm.kv + (int64_t)row * prm.ld + col0 : prm.zeros;
            bulk_g2s(A + SEP_A_TILE + lane * SEP_BST, src, row_bytes, &full[pf_slot]);
        } else if (lane == KRG_KD) {
            bulk_g2s(A, asrc, SEP_A_TILE * sizeof(double), &full[pf_slot]);
        }
#elif GSB_KRG_ISSUE == 2
        if (lane != 0) return;
        if (pf_round > 0) mbar_wait(&empty[pf_slot], (pf_round - 1) & 1);
        if (lead) {
            mbar_arrive_expect_tx(&full[pf_slot], tx);
            bulk_g2s(A, asrc, SEP_A_TILE * sizeof(double), &full[pf_slot]);
        }
#pragma unroll
        for (int d = 0; d < KRG_KD / SEP_WARPS; ++d) {
            const int lr = warp * (KRG_KD / SEP_WARPS) + d;
            const int row = pf.s * KRG_KD + lr;
            const double *src = row < prm.K ? prm.kv + (int64_t)row * prm.ld + col0 : prm.zeros;
            bulk_g2s(A + SEP_A_TILE + lr * SEP_BST, src, row_bytes, &full[pf_slot]);
        }
#else
        if (!lead || lane != 0) return;
        if (pf_round > 0) mbar_wait(&empty[pf_slot], (pf_round - 1) & 1);
        mbar_arrive_expect_tx(&full[pf_slot], tx);
        bulk_g2s(A, asrc, SEP_A_TILE * sizeof(double), &full[pf_slot]);
        const double *src = prm.kv + (int64_t)pf.s * KRG_KD * prm.ld + col0;
        const int nreal = min(KRG_KD, prm.K - pf.s * KRG_KD);
#pragma unroll
        for (int d = 0; d < KRG_KD; ++d)
            bulk_g2s(A + SEP_A_TILE + d * SEP_BST, d < nreal ? src + (int64_t)d * prm.ld : prm.zeros, row_bytes,
                     &full[pf_slot]);
#endif
    };
    auto pf_advance = [&]() {
        if (++pf_slot == KRG_STAGES) { pf_slot = 0; ++pf_round; }
        cur_advance(pf, gridDim.x);
    };
#pragma unroll
    for (int p = 0; p < DEPTH; ++p) {
        if (pf.unit < n_units) {
            pf_issue(warp == 0);
            pf_advance();
        }
    }

    // warp -> (32-row band wr, 64-column half wc).  In the 8 diagonal stages of a row tile the operand
    // is zero above the diagonal, so band wr only has work in the first 2*wr + 2 of them.  Sub-partition k
    // hosts warps k and k + 4: pairing bands (0, 3) and (1, 2) there gives every FP64 pipe the same
    // 10 of 16 band-stages, and the diagonal block costs 5 stage times instead of 8.
    const int wc = warp & 1;
    const int wr = ((warp >> 1) & 1) == 0 ? ((warp >> 2) ? 3 : 0) : ((warp >> 2) ? 2 : 1);
    const int g = lane >> 2;
    const int t = lane & 3;
    const int a_off = (wr * 32 + g) * SEP_AST + t;
    const int b_off = SEP_A_TILE + t * SEP_BST + wc * 64 + g;

    int slot = 0;
    uint32_t round = 0;
    int turn = 0;
    for (int64_t unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
        const int64_t c = unit / prm.n_pairs;
        const int p = (int)(unit % prm.n_pairs);
        const int64_t col0 = c * SEP_TN;
        double esum[8][2];
#pragma unroll
        for (int j = 0; j < 8; ++j) esum[j][0] = esum[j][1] = 0.0;

        const int n_half = (p == prm.R - 1 - p) ? 1 : 2;
        for (int half = 0; half < n_half; ++half) {
            const int r = half == 0 ? prm.R - 1 - p : p;
            const int S = min(KRG_SPT * (r + 1), prm.n_dstages);
            double acc[4][8][2];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

            for (int s = 0; s < S; ++s) {
                if (pf.unit < n_units) {
                    pf_issue(warp == turn);
                    pf_advance();
                }
                turn = (turn + 1) & (SEP_WARPS - 1);
                __syncwarp();
                mbar_wait(&full[slot], round & 1);   // also for an idle band: keeps the warps within one ring round
                const double *Sm = stage_base + slot * KRG_STAGE_DOUBLES;
                if (s < KRG_SPT * r + 2 * wr + 2)
#pragma unroll
                for (int k4 = 0; k4 < KRG_KD / 4; ++k4) {
                    double af[4], bf[8];
#pragma unroll
                    for (int i = 0; i < 4; ++i) af[i] = Sm[a_off + i * 8 * SEP_AST + 4 * k4];
#pragma unroll
                    for (int j = 0; j < 8; ++j) bf[j] = Sm[b_off + 4 * k4 * SEP_BST + j * 8];
#pragma unroll
                    for (int j = 0; j < 8; ++j)
#pragma unroll
                        for (int i = 0; i < 4; ++i) dmma_884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
                }
                // Epilogue folded into the diagonal stages.  acc[i][j][e] = Y[row, col] with row = 128r + 32wr +
                // 8i + g, col = col0 + 64wc + 8j + 2t + e.  Diagonal stage sd = s - 8r holds the depth rows
                // 128r + 16sd .. +15 in shared memory: for band wr = sd / 2 these are exactly the kv rows of its
                // accumulator rows i = 2(sd & 1), 2(sd & 1) + 1, which are final after this stage's DMMAs
                // (the operand is zero beyond the diagonal).  error += kv[row, col] * Y[row, col] straight from
                // the stage buffer: no global loads, and the other bands keep the DMMA pipe busy meanwhile.
                // (Row-major variant only: 72.4 -> 68.1 ms for K = 1001, n = 128^3.  With pre-tiled right-hand
                // sides the end-of-sub-tile loads below hit L2 in one burst and measured 1.5 % faster.)
                if (!TILED && s >= KRG_SPT * r && ((s - KRG_SPT * r) >> 1) == wr) {
                    const double *Bt = Sm + SEP_A_TILE + g * SEP_BST + wc * 64 + 2 * t;
                    if (((s - KRG_SPT * r) & 1) == 0) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const double2 v0 = *reinterpret_cast<const double2 *>(Bt + j * 8);
                            const double2 v1 = *reinterpret_cast<const double2 *>(Bt + 8 * SEP_BST + j * 8);
                            esum[j][0] = fma(v0.x, acc[0][j][0], esum[j][0]);
                            esum[j][1] = fma(v0.y, acc[0][j][1], esum[j][1]);
                            esum[j][0] = fma(v1.x, acc[1][j][0], esum[j][0]);
                            esum[j][1] = fma(v1.y, acc[1][j][1], esum[j][1]);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const double2 v0 = *reinterpret_cast<const double2 *>(Bt + j * 8);
                            const double2 v1 = *reinterpret_cast<const double2 *>(Bt + 8 * SEP_BST + j * 8);
                            esum[j][0] = fma(v0.x, acc[2][j][0], esum[j][0]);
                            esum[j][1] = fma(v0.y, acc[2][j][1], esum[j][1]);
                            esum[j][0] = fma(v1.x, acc[3][j][0], esum[j][0]);
                            esum[j][1] = fma(v1.y, acc[3][j][1], esum[j][1]);
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[slot]);
                if (++slot == KRG_STAGES) { slot = 0; ++round; }
            }

            // end of the sub-tile.  TILED: error += kv[row, col] * Y[row, col] with kv from the pre-tiled chunk
            // (L2 hits).  Both: row K of the operand is w = M^T cond, i.e. the field.
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int row = r * SEP_TM + wr * 32 + i * 8 + g;
                if (TILED && row < prm.K) {
                    const double *kr = prm.btile + ((c * prm.n_dstages + row / KRG_KD) * KRG_KD + row % KRG_KD) * SEP_BST +
                                       wc * 64 + 2 * t;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const double2 v = *reinterpret_cast<const double2 *>(kr + j * 8);
                        esum[j][0] = fma(v.x, acc[i][j][0], esum[j][0]);
                        esum[j][1] = fma(v.y, acc[i][j][1], esum[j][1]);
                    }
                }
                if (row == prm.K) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int64_t col = col0 + wc * 64 + j * 8 + 2 * t;
                        if (col < prm.n) prm.field[col] = acc[i][j][0];
                        if (col + 1 < prm.n) prm.field[col + 1] = acc[i][j][1];
                    }
                }
            }
        }

        // reduce over the 8 row owners of a warp (lanes with equal t), then over the 4 row bands
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                double v = esum[j][e];
                v += __shfl_xor_sync(0xffffffffu, v, 4);
                v += __shfl_xor_sync(0xffffffffu, v, 8);
                v += __shfl_xor_sync(0xffffffffu, v, 16);
                esum[j][e] = v;
            }
        __syncthreads();   // red[] of the previous unit has been read
        if (g == 0) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                red[wr][wc * 64 + j * 8 + 2 * t] = esum[j][0];
                red[wr][wc * 64 + j * 8 + 2 * t + 1] = esum[j][1];
            }
        }
        __syncthreads();
        if (tid < SEP_TN && col0 + tid < prm.n)
            prm.partial[(int64_t)p * prm.n + col0 + tid] = ((red[0][tid] + red[1][tid]) + red[2][tid]) + red[3][tid];
    }
}

__global__ void krige_finish_kernel(const double *__restrict__ partial, int n_pairs, int64_t n,
                                    double *__restrict__ error)
{
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        double s = 0.0;
        for (int p = 0; p < n_pairs; ++p) s += partial[(int64_t)p * n + k];
        error[k] = s;
    }
}

// field[k] = sum_j w[j] kv[j,k]: one thread per pair of points, 4 independent row streams, coalesced
__global__ void __launch_bounds__(256) krige_field_kernel(const double *__restrict__ w, const double *__restrict__ kv,
                                                          int64_t ld, int K, int64_t n, double *__restrict__ field)
{
    __shared__ double ws[1024];
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    for (int j0 = 0; j0 < K; j0 += 1024) {
        const int cnt = min(1024, K - j0);
        __syncthreads();
        for (int e = threadIdx.x; e < cnt; e += blockDim.x) ws[e] = w[j0 + e];
        __syncthreads();
        if (k < n) {
            const double *col = kv + (int64_t)j0 * ld + k;
            int j = 0;
            for (; j + 4 <= cnt; j += 4) {
                s0 = fma(ws[j], col[(int64_t)j * ld], s0);
                s1 = fma(ws[j + 1], col[(int64_t)(j + 1) * ld], s1);
                s2 = fma(ws[j + 2], col[(int64_t)(j + 2) * ld], s2);
                s3 = fma(ws[j + 3], col[(int64_t)(j + 3) * ld], s3);
            }
            for (; j < cnt; ++j) s0 = fma(ws[j], col[(int64_t)j * ld], s0);
        }
    }
    if (k < n) field[k] = (s0 + s1) + (s2 + s3);
}

// aligned repack for callers whose kv is not 16-byte aligned / has odd ld or n
__global__ void krige_repack_kernel(const double *__restrict__ src, int64_t src_ld, int K, int64_t n,
                                    double *__restrict__ dst, int64_t dst_ld)
{
    const int64_t total = (int64_t)K * dst_ld;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t j = e / dst_ld, k = e % dst_ld;
        dst[e] = k < n ? src[j * src_ld + k] : 0.0;
    }
}

inline int launch_krige(const KrigeParams &kp, bool tiled, int sm_count, cudaStream_t st)
{
    static std::atomic<uint64_t> attr_set{0};
    if (first_launch_on_device(attr_set)) {
        GSB_CUDA(cudaFuncSetAttribute(krige_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)KRG_SMEM_BYTES));
        GSB_CUDA(cudaFuncSetAttribute(krige_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)KRG_SMEM_BYTES));
    }
    const int64_t n_units = kp.n_col_tiles * kp.n_pairs;
    dim3 grid((unsigned)std::min<int64_t>(n_units, sm_count));
    if (tiled) krige_kernel<true><<<grid, SEP_THREADS, KRG_SMEM_BYTES, st>>>(kp);
    else krige_kernel<false><<<grid, SEP_THREADS, KRG_SMEM_BYTES, st>>>(kp);
    g_launches.fetch_add(1);
    GSB_CUDA(cudaGetLastError());
    return GSB_OK;
}

// ---------------------------------------------------------------------------------------------
// right-hand sides on the device (reference: Krige._get_krige_vecs, krige/base.py:359-388):
//   rows 0..C-1      cov(|x_cond_i - x_k|)  with cdist on the isometrised positions (base.py:430-450) and
//                    CovModel.covariance / cov_nugget (covmodel/tools.py:65-76, covmodel/base.py:313-320)
//   row  C           1 when the system is unbiased (base.py:377-378)
//   rows after that  drift rows supplied by the caller (functional and external drift, base.py:379-387)
// ---------------------------------------------------------------------------------------------
struct CovParams {
    int type;             // GSB_COV_*
    int exact;            // cov_nugget instead of covariance (Krige.exact)
    double var, len_rescaled, sill, param;
};

// normalised correlation cor(h) of the supported models (src/gstools/covmodel/models.py)
__device__ __forceinline__ double cov_cor(const CovParams &m, double h)
{
    switch (m.type) {
    case GSB_COV_GAUSSIAN: return exp(-(h * h));                                        // models.py:139-141
    case GSB_COV_EXPONENTIAL: return exp(-h);                                           // models.py:213-215
    case GSB_COV_STABLE: return exp(-pow(h, m.param));                                  // models.py:343-345
    case GSB_COV_RATIONAL: return pow(1.0 + h * h / m.param, -m.param);                 // models.py:610-612
    case GSB_COV_CUBIC: {                                                               // models.py:654-657
        const double x = fmin(h, 1.0), x2 = x * x, x3 = x2 * x;
        return (((1.0 - 7.0 * x2) + 8.75 * x3) - 3.5 * (x3 * x2)) + 0.75 * (x3 * x2 * x2);
    }
    case GSB_COV_LINEAR: return fmax(1.0 - h, 0.0);                                     // models.py:687-689
    case GSB_COV_CIRCULAR:                                                              // models.py:728-738
        return h < 1.0 ? 2.0 / 3.141592653589793 * (acos(h) - h * sqrt(1.0 - h * h)) : 0.0;
    case GSB_COV_SPHERICAL: {                                                           // models.py:774-777
        const double x = fmin(h, 1.0);
        return (1.0 - 1.5 * x) + 0.5 * (x * x * x);
    }
    default: return 0.0;
    }
}

__device__ __forceinline__ double cov_value(const CovParams &m, double r)
{
    // cov_nugget: np.isclose(r, 0) <=> |r| <= 1e-8 -> sill (covmodel/base.py:313-320)
    if (m.exact && r <= 1e-8) return m.sill;
    return m.var * cov_cor(m, r / m.len_rescaled);
}

struct KvgenParams {
    CovParams cov;
    int dim;
    int C;                 // conditioning points (covariance rows)
    int K;                 // kriging system size
    int unbiased;
    int n_dstages;
    const double *cond_pos;    // (dim, C) isometrised
    // evaluation points of this column chunk: flat (dim, n) with row stride pos_ld, or a structured mesh
    const double *pos;
    int64_t pos_ld;
    const double *axes;        // structured: concatenated axes + matrix (as ExpandParams)
    int64_t axis_off[GSB_MAX_DIM];
    int64_t axis_len[GSB_MAX_DIM];
    double matrix[GSB_MAX_DIM * GSB_MAX_DIM];
    int64_t col_begin;         // first point of this chunk (global index)
    int64_t n;                 // points in this chunk
    const double *tail;        // (K - C - unbiased, n_total) drift rows, row stride tail_ld; indexed globally
    int64_t tail_ld;
    double *btile;             // out: pre-tiled chunk (TILED layout)
    const double *w;           // field-only kernel: M^T cond
    double *field;
};

template <int D>
__device__ __forceinline__ void kvgen_point(const KvgenParams &prm, int64_t gcol, double (&x)[D])
{
    if (prm.pos) {
#pragma unroll
        for (int t = 0; t < D; ++t) x[t] = prm.pos[t * prm.pos_ld + gcol];
    } else {   // generate_grid + isometrize on the fly (geometric.py:340-356, covmodel/base.py:572-582)
        double gpt[D];
        int64_t rem = gcol;
#pragma unroll
        for (int t = D - 1; t >= 0; --t) {
            const int64_t it = rem % prm.axis_len[t];
            rem /= prm.axis_len[t];
            gpt[t] = prm.axes[prm.axis_off[t] + it];
        }
#pragma unroll
        for (int t = 0; t < D; ++t) {
            double v = 0.0;
#pragma unroll
            for (int u = 0; u < D; ++u) v += prm.matrix[t * D + u] * gpt[u];
            x[t] = v;
        }
    }
}

template <int D>
__device__ __forceinline__ double kvgen_entry(const KvgenParams &prm, int row, int64_t gcol, const double (&x)[D],
                                              const double *cpos /* [D] of this row, or nullptr */)
{
    if (row < prm.C) {
        double s2 = 0.0;
#pragma unroll
        for (int t = 0; t < D; ++t) {
            const double dlt = __dsub_rn(cpos[t], x[t]);
            s2 = __dadd_rn(s2, __dmul_rn(dlt, dlt));     // scipy cdist: plain sum of squares, then sqrt
        }
        return cov_value(prm.cov, sqrt(s2));
    }
    if (row < prm.C + prm.unbiased) return 1.0;
    if (row < prm.K) return prm.tail[(int64_t)(row - prm.C - prm.unbiased) * prm.tail_ld + gcol];
    return 0.0;
}

// cor(h) with the model fixed at compile time (same expressions as cov_cor)
template <int KIND>
__device__ __forceinline__ double cov_cor_t(double param, double h)
{
    CovParams m;
    m.type = KIND;
    m.param = param;
    return cov_cor(m, h);     // the switch folds away: KIND is a constant
}

// covariance rows rr = half, half + 2, ... < n_cov of the staged block (rows r0 + rr < C)
template <int D, int KIND>
__device__ __forceinline__ void kvgen_cov_rows(const KvgenParams &prm, const double (&x)[D], const double *cp,
                                               int half, int n_cov, int r0, int cl, double *T)
{
    const double var = prm.cov.var, len = prm.cov.len_rescaled, sill = prm.cov.sill, param = prm.cov.param;
    const bool exact = prm.cov.exact != 0;
#pragma unroll 1
    for (int rr = half; rr < n_cov; rr += 2) {
        const double *c = cp + rr * D;
        double s2 = 0.0;
#pragma unroll
        for (int t = 0; t < D; ++t) {
            const double dlt = __dsub_rn(c[t], x[t]);
            s2 = __dadd_rn(s2, __dmul_rn(dlt, dlt));     // scipy cdist: plain sum of squares, then sqrt
        }
        const double r = sqrt(s2);
        double v = var * cov_cor_t<KIND>(param, r / len);
        if (exact && r <= 1e-8) v = sill;                  // cov_nugget (covmodel/base.py:313-320)
        const int row = r0 + rr;
        T[(row / KRG_KD) * SEP_B_TILE + (row % KRG_KD) * SEP_BST + cl] = v;
    }
}

// one CTA per column tile of 128 points: every thread owns one point and every second row, walks down
// all depth stages (position and constants set up once) and writes in the stage layout
constexpr int KVGEN_ROWS = 1024;     // conditioning positions staged in shared memory per pass
template <int D>
__global__ void __launch_bounds__(256) kvgen_kernel(const KvgenParams prm)
{
    __shared__ double cp[KVGEN_ROWS][D];
    const int64_t ct = blockIdx.x;
    const int cl = threadIdx.x % SEP_TN;
    const int half = threadIdx.x / SEP_TN;
    const int64_t lcol = ct * SEP_TN + cl;
    const int64_t gcol = prm.col_begin + lcol;
    double *T = prm.btile + ct * prm.n_dstages * SEP_B_TILE;
    double x[D];
    const bool live = lcol < prm.n;
    if (live) kvgen_point<D>(prm, gcol, x);
    // gridDim.y CTAs share the depth stages of a column tile (finer grain: fewer idle SMs in the last wave)
    const int st_per = (prm.n_dstages + gridDim.y - 1) / gridDim.y;
    const int row_lo = min(prm.n_dstages, (int)blockIdx.y * st_per) * KRG_KD;
    const int n_rows = min(prm.n_dstages, ((int)blockIdx.y + 1) * st_per) * KRG_KD;
    for (int r0 = row_lo; r0 < n_rows; r0 += KVGEN_ROWS) {
        const int cnt = min(KVGEN_ROWS, n_rows - r0);
        __syncthreads();
        for (int e = threadIdx.x; e < cnt * D; e += blockDim.x) {
            const int rr = e / D, t = e % D, row = r0 + rr;
            cp[rr][t] = row < prm.C ? prm.cond_pos[(int64_t)t * prm.C + row] : 0.0;
        }
        __syncthreads();
        // covariance rows first, with the model's cor(h) resolved at compile time (no per-element
        // switch, no per-element row classification), then the few non-covariance rows generically
        const int n_cov = live ? max(0, min(cnt, prm.C - r0)) : 0;
        switch (prm.cov.type) {
#define GSB_KVGEN_CASE(KIND) \
        case KIND: kvgen_cov_rows<D, KIND>(prm, x, &cp[0][0], half, n_cov, r0, cl, T); break;
            GSB_KVGEN_CASE(GSB_COV_GAUSSIAN) GSB_KVGEN_CASE(GSB_COV_EXPONENTIAL) GSB_KVGEN_CASE(GSB_COV_STABLE)
            GSB_KVGEN_CASE(GSB_COV_RATIONAL) GSB_KVGEN_CASE(GSB_COV_CUBIC) GSB_KVGEN_CASE(GSB_COV_LINEAR)
            GSB_KVGEN_CASE(GSB_COV_CIRCULAR) GSB_KVGEN_CASE(GSB_COV_SPHERICAL)
#undef GSB_KVGEN_CASE
        default: break;
        }
        // first row of this thread's parity at or after n_cov
        for (int rr = n_cov + ((n_cov ^ half) & 1); rr < cnt; rr += 2) {
            const int row = r0 + rr;
            T[(row / KRG_KD) * SEP_B_TILE + (row % KRG_KD) * SEP_BST + cl] =
                live ? kvgen_entry<D>(prm, row, gcol, x, cp[rr]) : 0.0;
        }
    }
    // the 4 padding doubles of every row
    for (int e = threadIdx.x; e < (n_rows - row_lo) * (SEP_BST - SEP_TN); e += blockDim.x) {
        const int row = row_lo + e / (SEP_BST - SEP_TN), q = e % (SEP_BST - SEP_TN);
        T[(row / KRG_KD) * SEP_B_TILE + (row % KRG_KD) * SEP_BST + SEP_TN + q] = 0.0;
    }
}

// field only (return_var = False): field[k] = sum_row w[row] kv[row, k] with kv generated on the fly.
// The conditioning positions and w are staged through shared memory in passes of FIELD_GEN_ROWS rows
// (an even number, so the even / odd partial sums see the same rows whatever the system size): no limit
// on the size of the kriging system.
constexpr int FIELD_GEN_ROWS = 1024;

template <int D>
__global__ void __launch_bounds__(256) krige_field_gen_kernel(const KvgenParams prm)
{
    __shared__ double cps[FIELD_GEN_ROWS * D];   // [row][D] conditioning positions of this pass
    __shared__ double ws[FIELD_GEN_ROWS];
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const bool live = k < prm.n;
    double x[D];
    if (live) kvgen_point<D>(prm, prm.col_begin + k, x);
    double s0 = 0.0, s1 = 0.0;
    for (int r0 = 0; r0 < prm.K; r0 += FIELD_GEN_ROWS) {
        const int cnt = min(FIELD_GEN_ROWS, prm.K - r0);
        __syncthreads();   // the previous pass has been consumed
        for (int e = threadIdx.x; e < cnt * D; e += blockDim.x) {
            const int rr = e / D, t = e % D, row = r0 + rr;
            cps[e] = row < prm.C ? prm.cond_pos[(int64_t)t * prm.C + row] : 0.0;
        }
        for (int e = threadIdx.x; e < cnt; e += blockDim.x) ws[e] = prm.w[r0 + e];
        __syncthreads();
        if (!live) continue;
        int rr = 0;
        for (; rr + 2 <= cnt; rr += 2) {
            s0 = fma(ws[rr], kvgen_entry<D>(prm, r0 + rr, prm.col_begin + k, x, cps + (size_t)rr * D), s0);
            s1 = fma(ws[rr + 1], kvgen_entry<D>(prm, r0 + rr + 1, prm.col_begin + k, x, cps + (size_t)(rr + 1) * D), s1);
        }
        if (rr < cnt) s0 = fma(ws[rr], kvgen_entry<D>(prm, r0 + rr, prm.col_begin + k, x, cps + (size_t)rr * D), s0);
    }
    if (live) prm.field[k] = s0 + s1;
}

}  // namespace gsb
