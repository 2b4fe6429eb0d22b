// gsb_separable.cuh -- separable summation on structured (rectilinear) meshes for sm_100a.
//
// On a structured mesh the position of grid node (i_0, ..., i_{d-1}) is M (a_0[i_0], ...,
// a_{d-1}[i_{d-1}]) (reference: src/gstools/tools/geometric.py:340-356 generate_grid, then
// src/gstools/covmodel/base.py:572-582 isometrize).  With k' = M^T k the phase splits per axis,
//     k . x = sum_t k'_t a_t[i_t],
// and the sum of the reference summator (src/gstools/field/generator.py:193-199) becomes
//     u[row, c] = sum_j Re( A_j(row) * E_j(c) ) = sum_j  Ar[row,j] Cz[j,c] + (-Ai[row,j]) Sz[j,c]
// with  A_j(row) = (z1_j - i z2_j) prod_{t<d-1} exp(i k'_{t,j} a_t[i_t])   (row = all axes but the last)
//       E_j(c)   = exp(i k'_{d-1,j} a_{d-1}[c]) = Cz + i Sz               (c = last, contiguous axis).
// That is a real fp64 contraction of depth 2N costing 2 DFMA per (point, mode) instead of the
// ~D+13 of the direct kernel: all sin/cos work moves into per-axis tables of size (len_t x N)
// built once per call with full-accuracy sincos.
//
// Three kernels:
//   1. build_tables_kernel   per-axis phase tables; the last-axis table is written PRE-TILED in the
//                            exact shared-memory layout of a pipeline stage (one bulk copy each).
//                            For the incompressible variant (generator.py:479-495) the projector
//                            p_t(k_j) is folded into the last-axis table of component t.
//   2. agen_kernel           A operand: one complex product per (row, mode), written pre-tiled
//                            (memory bound, ~5 % of the contraction time, runs on a helper stream
//                            concurrently with the contraction of the previous row chunk).
//   3. contract_kernel       one CTA of 8 warps per 128x128 output tile, 1 CTA per SM:
//        TMA                  two cp.async.bulk copies per stage of KC modes (a 20 KB A tile and a
//                             16.5 KB B tile) completing on the stage's "full" mbarrier, issued
//                             three stages ahead by lane 0 of a warp that rotates with the stage
//                             (it first checks the slot's "empty" mbarrier);
//        every warp           32x64 warp tile = 4x8 DMMA.8x8x4 accumulator tiles (64 fp64
//                             accumulators per thread); per 2 modes 4 + 8 fragment doubles by
//                             conflict-free LDS.64 and 32 DMMA; modes in ascending order
//                             (deterministic, no atomics); registers -> global in the epilogue.
//      An earlier single-kernel version generated A inside the contraction kernel; ncu and the
//      bisect microbenchmark (profiles/) showed the FP64 products, their loads and stores costing
//      ~12 % of the DMMA issue slots however they were placed, while this split reaches 97 %.
//
// Why DMMA and not DFMA: B200 runs DMMA.8x8x4 at the same FMA rate as DFMA (measured 18.5 vs
// 18.4 TFMA/s), but one instruction carries 256 FMAs with 8 register reads.  A register-tiled DFMA
// version of the contraction topped out at 73 % of the FP64 peak: 3 x 64-bit operands per DFMA
// exceed what the register file sustains once the operand-reuse cache is lost between warps.
#pragma once

#include <algorithm>
#include <type_traits>

#include "gsb_common.cuh"

namespace gsb {

constexpr int SEP_TM = 128;      // rows per CTA tile
constexpr int SEP_TN = 128;      // columns per CTA tile
#ifndef GSB_SEP_KC
#define GSB_SEP_KC 8
#endif
#ifndef GSB_SEP_STAGES
#define GSB_SEP_STAGES 4
#endif
constexpr int SEP_KC = GSB_SEP_KC;        // modes per pipeline stage (multiple of 2)
constexpr int SEP_STAGES = GSB_SEP_STAGES;
static_assert(SEP_KC % 2 == 0 && SEP_STAGES >= 3, "pipeline shape");
static_assert((8 & (8 - 1)) == 0, "warp rotation uses a power-of-two mask");
constexpr int SEP_MAX_ROW_AXES = GSB_MAX_DIM - 1;
constexpr int SEP_WARPS = 8;
constexpr int SEP_THREADS = SEP_WARPS * 32;

// Stage layout in shared memory == tile layout in global memory (doubles):
//   A tile [SEP_TM rows][SEP_AST]   k = 2*kc + part (part 0: Re A, part 1: -Im A); the row stride
//                                   2*KC+4 makes the 8x4 fragment loads (LDS.64) hit 32 distinct
//                                   banks per half warp
//   B tile [2*KC][SEP_BST]          row 2*kc + part (part 0: p*cos, part 1: p*sin), stride 132
constexpr int SEP_AST = 2 * SEP_KC + 4;
constexpr int SEP_BST = SEP_TN + 4;
constexpr int SEP_A_TILE = SEP_TM * SEP_AST;          // doubles
constexpr int SEP_B_TILE = 2 * SEP_KC * SEP_BST;      // doubles
constexpr int SEP_C_BLOCK = 2 * SEP_KC;               // doubles: KC complex slow-axis factors (scaled variant)
constexpr int SEP_STAGE_DOUBLES = SEP_A_TILE + SEP_B_TILE + SEP_C_BLOCK;
static_assert((SEP_A_TILE * 8) % 16 == 0 && (SEP_B_TILE * 8) % 16 == 0, "bulk copy granularity");
static_assert((SEP_A_TILE / 2) % SEP_TM == 0, "agen copy-out loop");
constexpr size_t SEP_SMEM_BYTES =
    (size_t)SEP_STAGES * SEP_STAGE_DOUBLES * sizeof(double) + 2 * SEP_STAGES * sizeof(uint64_t) + 128;

// ---------------------------------------------------------------------------------------------
// 1. table builder.  Full-accuracy sincos (libdevice), cost O((sum_t len_t) N): negligible.
// ---------------------------------------------------------------------------------------------
struct TableParams {
    const double *cov;     // (n_batch, dim, n_modes)
    const double *z1;      // (n_batch, n_modes)
    const double *z2;
    const double *sf;      // optional per-mode spectrum factor (Fourier generator), (n_batch, n_modes)
    const double *axes;    // concatenated axis coordinates
    int64_t axis_off[GSB_MAX_DIM];
    int64_t axis_len[GSB_MAX_DIM];
    double matrix[GSB_MAX_DIM * GSB_MAX_DIM];  // row-major (dim x dim) isometrisation matrix
    int dim;               // dimension of the wave vectors / the matrix
    int n_axes;            // table axes: dim, or dim - 1 when the last two mesh axes are folded into one
    int64_t fold_len;      // folded: length of the real last axis (0 = not folded); table entry i of the
    int64_t fold_off;      //   last table axis is node (i / fold_len, i % fold_len) of the last two axes
    int64_t n_modes;
    int n_modes_pad;       // multiple of SEP_KC; padded modes are zero
    int ncomp;             // 1 scalar, dim for the incompressible field
    // row-axis tables: erow[t][b][j * len_t + i] = exp(i k'_{t,j} a_t[i]) as (cos, sin); t = 0 is
    // pre-multiplied by (z1_j - i z2_j)
    double2 *erow[SEP_MAX_ROW_AXES];
    int64_t erow_bstride[SEP_MAX_ROW_AXES];
    // last-axis table, pre-tiled: btile[((b*ncomp + comp)*n_col_tiles + ct)*n_stages + s] is one
    // SEP_B_TILE block
    double *btile;
    int n_col_tiles;
    // scaled variant only: the LAST row axis pre-tiled like an A tile, raw (cos, sin):
    // ytab[((b*n_ytiles + yt)*n_stages + s)] is one SEP_A_TILE block (rows beyond the axis are zero)
    double *ytab;
    int n_ytiles;
};

__global__ void build_tables_kernel(const TableParams tp)
{
    const int t = blockIdx.y;                 // axis
    const int64_t b = blockIdx.z;             // batch entry
    const int64_t len = tp.axis_len[t];
    const bool last = (t == tp.n_axes - 1);
    const bool ytile = (tp.ytab != nullptr && t == tp.n_axes - 2);
    const int64_t width = last ? (int64_t)tp.n_col_tiles * SEP_TN : (ytile ? (int64_t)tp.n_ytiles * SEP_TM : len);
    const int64_t total = width * tp.n_modes_pad;
    const int n_stages = tp.n_modes_pad / SEP_KC;
    const double *cov = tp.cov + b * tp.dim * tp.n_modes;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t j = idx / width;
        const int64_t i = idx - j * width;
        double c = 0.0, s = 0.0;
        if (j < tp.n_modes && i < len) {
            // k'_t = sum_s M[s][t] k_s   (phase = k . (M a) = (M^T k) . a)
            double kp = 0.0;
            for (int s2 = 0; s2 < tp.dim; ++s2)
                kp = fma(tp.matrix[s2 * tp.dim + t], cov[(int64_t)s2 * tp.n_modes + j], kp);
            double phase;
            if (last && tp.fold_len > 0) {   // folded trailing axes: k'_y y + k'_z z
                double kz = 0.0;
                for (int s2 = 0; s2 < tp.dim; ++s2)
                    kz = fma(tp.matrix[s2 * tp.dim + t + 1], cov[(int64_t)s2 * tp.n_modes + j], kz);
                phase = fma(kp, tp.axes[tp.axis_off[t] + i / tp.fold_len], kz * tp.axes[tp.fold_off + i % tp.fold_len]);
            } else {
                phase = kp * tp.axes[tp.axis_off[t] + i];
            }
            sincos(phase, &s, &c);
            if (t == 0 && !last) {  // fold the complex weight (z1 - i z2) into the first row axis
                const double w = tp.sf ? tp.sf[b * tp.n_modes + j] : 1.0;
                const double a = w * tp.z1[b * tp.n_modes + j], bb = w * tp.z2[b * tp.n_modes + j];
                const double re = a * c + bb * s;
                const double im = a * s - bb * c;
                c = re;
                s = im;
            }
        }
        if (!last) {
            if (i < len) tp.erow[t][b * tp.erow_bstride[t] + j * len + i] = make_double2(c, s);
            if (ytile) {
                const int yt = (int)(i / SEP_TM), r = (int)(i % SEP_TM);
                double *tile = tp.ytab + (((b * tp.n_ytiles + yt) * n_stages + (j / SEP_KC)) * (int64_t)SEP_A_TILE);
                *reinterpret_cast<double2 *>(tile + r * SEP_AST + 2 * (j % SEP_KC)) = make_double2(c, s);
            }
            continue;
        }
        const int ct = (int)(i / SEP_TN), col = (int)(i % SEP_TN);
        const int st = (int)(j / SEP_KC), kc = (int)(j % SEP_KC);
        double k2 = 0.0, k0 = 0.0;
        if (tp.ncomp > 1 && j < tp.n_modes) {
            for (int s2 = 0; s2 < tp.dim; ++s2) {
                const double k = cov[(int64_t)s2 * tp.n_modes + j];
                k2 += k * k;
            }
            k0 = cov[j];
        }
        for (int comp = 0; comp < tp.ncomp; ++comp) {
            double p = 1.0;
            if (tp.ncomp > 1) {
                // incompressible projector on the ORIGINAL wave vector (generator.py:479-495)
                p = 0.0;
                if (j < tp.n_modes) {
                    const double e = (comp == 0) ? 1.0 : 0.0;
                    p = e - cov[(int64_t)comp * tp.n_modes + j] * k0 / k2;
                }
            }
            double *tile = tp.btile +
                           ((((b * tp.ncomp + comp) * tp.n_col_tiles + ct) * n_stages + st) * (int64_t)SEP_B_TILE);
            tile[(2 * kc) * SEP_BST + col] = p * c;
            tile[(2 * kc + 1) * SEP_BST + col] = p * s;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// 2. A operand generator: atile[((f*n_row_tiles + rt)*n_stages + s)] is one SEP_A_TILE block holding
//    rows rt*128 .. rt*128+127 of field-batch entry f for modes s*KC .. s*KC+KC-1.
// ---------------------------------------------------------------------------------------------
struct AgenParams {
    const double2 *erow[SEP_MAX_ROW_AXES];
    int64_t erow_bstride[SEP_MAX_ROW_AXES];
    int64_t row_len[SEP_MAX_ROW_AXES];   // full extent of every row axis
    int n_row_axes;
    int64_t n_rows;        // prod(row_len)
    int64_t row_begin;     // first row of this chunk (multiple of SEP_TM)
    int n_row_tiles;       // row tiles in this chunk
    int n_modes_pad;
    int64_t batch0;        // first batch entry of this chunk
    double *atile;
};

template <int NRA>
__global__ void __launch_bounds__(SEP_TM) agen_kernel(const AgenParams ap)
{
    __shared__ __align__(16) double tile[SEP_A_TILE];   // staged so the global stores are coalesced
    const int rt = blockIdx.x;
    const int64_t f = blockIdx.z;
    const int n_stages = ap.n_modes_pad / SEP_KC;
    const int row = threadIdx.x;
    int64_t r = ap.row_begin + (int64_t)rt * SEP_TM + row;
    if (r >= ap.n_rows) r = ap.n_rows - 1;   // clamp; those rows are never stored by the contraction
    const double2 *ep[NRA];
#pragma unroll
    for (int t = NRA - 1; t >= 0; --t) {
        const int64_t it = r % ap.row_len[t];
        r /= ap.row_len[t];
        ep[t] = ap.erow[t] + (ap.batch0 + f) * ap.erow_bstride[t] + it;
    }
    double *tiles = ap.atile + (((int64_t)f * ap.n_row_tiles + rt) * n_stages) * (int64_t)SEP_A_TILE;
    // the pad columns 2*KC .. AST-1 are never read by the contraction; keep them defined
    tile[row * SEP_AST + 2 * SEP_KC + 0] = 0.0;
    tile[row * SEP_AST + 2 * SEP_KC + 1] = 0.0;
    tile[row * SEP_AST + 2 * SEP_KC + 2] = 0.0;
    tile[row * SEP_AST + 2 * SEP_KC + 3] = 0.0;
    // blockIdx.y strides over the stages so that small meshes still fill the GPU
    for (int s = blockIdx.y; s < n_stages; s += gridDim.y) {
        double2 e[SEP_KC];
#pragma unroll
        for (int u = 0; u < SEP_KC; ++u) e[u] = __ldg(ep[0] + ((int64_t)s * SEP_KC + u) * ap.row_len[0]);
#pragma unroll
        for (int t = 1; t < NRA; ++t) {
#pragma unroll
            for (int u = 0; u < SEP_KC; ++u) {
                const double2 f2 = __ldg(ep[t] + ((int64_t)s * SEP_KC + u) * ap.row_len[t]);
                const double re = e[u].x * f2.x - e[u].y * f2.y;
                const double im = e[u].x * f2.y + e[u].y * f2.x;
                e[u].x = re;
                e[u].y = im;
            }
        }
        __syncthreads();   // previous stage's copy-out is complete
        double2 *dst = reinterpret_cast<double2 *>(tile + row * SEP_AST);
#pragma unroll
        for (int u = 0; u < SEP_KC; ++u) dst[u] = make_double2(e[u].x, -e[u].y);
        __syncthreads();
        const double2 *src = reinterpret_cast<const double2 *>(tile);
        double2 *gdst = reinterpret_cast<double2 *>(tiles + (int64_t)s * SEP_A_TILE);
#pragma unroll
        for (int i = 0; i < SEP_A_TILE / 2 / SEP_TM; ++i) gdst[i * SEP_TM + row] = src[i * SEP_TM + row];
    }
}

// Scaled variant: slow-axis phase factors c[b][s][j] = prod_{t < NRA-1} E_t[j][i_t(s)] (weights folded
// in through axis 0), s = flattened index over all row axes but the last.
struct CtabParams {
    const double2 *erow[SEP_MAX_ROW_AXES];
    int64_t erow_bstride[SEP_MAX_ROW_AXES];
    int64_t row_len[SEP_MAX_ROW_AXES];
    int n_slow_axes;       // NRA - 1 >= 1
    int64_t n_slow;        // prod(row_len[:n_slow_axes])
    int n_modes_pad;
    double2 *ctab;         // (n_batch, n_slow, n_modes_pad)
};

__global__ void ctab_kernel(const CtabParams cp)
{
    const int64_t b = blockIdx.z;
    const int64_t total = cp.n_slow * cp.n_modes_pad;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t sidx = idx / cp.n_modes_pad;
        const int64_t j = idx - sidx * cp.n_modes_pad;
        int64_t r = sidx;
        double2 e = make_double2(1.0, 0.0);
        for (int t = cp.n_slow_axes - 1; t >= 0; --t) {
            const int64_t it = r % cp.row_len[t];
            r /= cp.row_len[t];
            const double2 f = cp.erow[t][b * cp.erow_bstride[t] + j * cp.row_len[t] + it];
            const double re = e.x * f.x - e.y * f.y;
            const double im = e.x * f.y + e.y * f.x;
            e.x = re;
            e.y = im;
        }
        cp.ctab[(b * cp.n_slow + sidx) * cp.n_modes_pad + j] = e;
    }
}

// ---------------------------------------------------------------------------------------------
// 3. the contraction
// ---------------------------------------------------------------------------------------------
struct ContractParams {
    const double *atile;   // chunk-local A tiles, see AgenParams
    const double *btile;   // whole-call B tiles, see TableParams
    int n_row_tiles;       // row tiles in this chunk (gridDim.y)
    int n_col_tiles;       // gridDim.x
    int n_stages;
    int ncomp;
    int64_t n_fields;      // (batch entries in this chunk) * ncomp
    int64_t batch0;        // first batch entry of this chunk
    int64_t n_rows;        // rows of one field
    int64_t row_begin;     // first row of this chunk
    int64_t lc;            // length of the last axis
    double *out;           // field (batch, comp) starts at out + (batch*ncomp + comp)*out_fstride
    int64_t out_fstride;
    // scaled variant (template SCALED): A tiles come from the pre-tiled last row axis and are
    // multiplied by the slow-axis phase factor in the consumer; row tile rt = s_slow*n_ytiles + yt
    const double *ytab;
    const double2 *ctab;
    int n_ytiles;
    int64_t n_slow;
    int64_t ly;            // length of the last row axis
    int n_modes_pad;
    int64_t rt0;           // first (global) row tile of this launch
    Epi epi;               // fused caller epilogue (off: raw sums)
    int no_partial;        // tuning: 1 = always the full-tile kernel (option "partial_tiles" = 0)
};

__device__ __forceinline__ void dmma_884(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// SCALED = false: A tiles pre-generated by agen_kernel (best when an A tile feeds >= 3 output tiles).
// SCALED = true : no A generation at all.  The A tile is the pre-tiled table of the LAST row axis
//                 (shared by all slow indices, L2 resident); the phase factor of the slower axes,
//                 c_j(s), arrives as a third 128-byte bulk copy per stage and each consumer rescales
//                 its own A fragments (2 FP64 ops per fragment).  Costs ~9 % of the DMMA rate
//                 (profiles/r01_microbench_dmma_tma.log) but no scratch, no extra HBM traffic, and
//                 it does not depend on how often an A tile is reused.
//
// PARTIAL = true : the last column tile of every row is narrower than 128 (last axis not a multiple of 128):
//                  warps skip the 8-column groups beyond the mesh.  A separate instantiation, so that the
//                  full-tile kernel keeps its code and register count.
template <bool SCALED, bool PARTIAL = false>
__global__ void __launch_bounds__(SEP_THREADS, 1) contract_kernel(const ContractParams prm)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *stage_base = reinterpret_cast<double *>(smem_raw);
    uint64_t *full = reinterpret_cast<uint64_t *>(stage_base + SEP_STAGES * SEP_STAGE_DOUBLES);
    uint64_t *empty = full + SEP_STAGES;

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    const int n_stages = prm.n_stages;
    // PERSISTENT: one CTA per SM walks over the output tiles blockIdx.x, blockIdx.x + gridDim.x, ...
    // (column tile fastest, so the CTAs working side by side share their A tiles through L2).
    // The CTA never leaves its SM, so the concurrently running A-generation kernel can only take
    // the resources this kernel leaves free and truly overlaps with it.
    const int64_t n_tiles = (int64_t)prm.n_col_tiles * prm.n_row_tiles * prm.n_fields;

    if (tid == 0) {
        for (int s = 0; s < SEP_STAGES; ++s) {
            mbar_init(&full[s], 1);              // the issuing thread's arrive.expect_tx
            mbar_init(&empty[s], SEP_WARPS);     // one arrive per warp
        }
        fence_barrier_init();
    }
    __syncthreads();

    // ---- TMA: two bulk copies per pipeline step (A tile 20 KB, B tile 16.5 KB), issued by lane 0
    // of a warp that rotates with the step.  There is no dedicated producer warp: registers are
    // allocated to a CTA in units of four warps, so a ninth warp would cost as much as twelve.
    // Every thread tracks the prefetch cursor (tile, stage) incrementally -- no divisions in the
    // steady state.
    // tile -> column tile.  Column tile fastest, so that the CTAs working side by side share their A tiles.
    // With partial last column tiles the tiles of a row differ in cost; when the number of column tiles
    // divides the grid size every CTA would always get the same column tile (148 CTAs, 2 column tiles: half
    // of the CTAs only ever see the cheap partial tile), so the column index is rotated by the CTA's round.
    const bool skew = PARTIAL && (gridDim.x % prm.n_col_tiles) == 0;
    auto col_tile_of = [&](int64_t tile_idx) {
        const int64_t c0 = tile_idx % prm.n_col_tiles;
        return (int)(skew ? (c0 + tile_idx / gridDim.x) % prm.n_col_tiles : c0);
    };
    constexpr int DEPTH = SEP_STAGES - 2;   // steps in flight ahead of the one being contracted
    int64_t pf_tile = blockIdx.x;
    int pf_s = 0;
    int pf_slot = 0;
    uint32_t pf_round = 0;                  // how often pf_slot has wrapped
    const double *pf_a = nullptr, *pf_b = nullptr, *pf_c = nullptr;
    auto pf_decode = [&]() {
        const int ct = col_tile_of(pf_tile);
        const int64_t rest = pf_tile / prm.n_col_tiles;
        const int rt = (int)(rest % prm.n_row_tiles);
        const int64_t z = rest / prm.n_row_tiles;          // field-local index * ncomp + comp
        const int comp = (int)(z % prm.ncomp);
        const int64_t fl = z / prm.ncomp;
        if (SCALED) {
            const int64_t sl = (prm.rt0 + rt) / prm.n_ytiles;
            const int yt = (int)((prm.rt0 + rt) % prm.n_ytiles);
            pf_a = prm.ytab + (((prm.batch0 + fl) * prm.n_ytiles + yt) * (int64_t)n_stages) * SEP_A_TILE;
            pf_c = reinterpret_cast<const double *>(prm.ctab + ((prm.batch0 + fl) * prm.n_slow + sl) * prm.n_modes_pad);
        } else {
            pf_a = prm.atile + ((fl * prm.n_row_tiles + rt) * (int64_t)n_stages) * SEP_A_TILE;
        }
        pf_b = prm.btile +
               ((((prm.batch0 + fl) * prm.ncomp + comp) * prm.n_col_tiles + ct) * (int64_t)n_stages) * SEP_B_TILE;
    };
    auto pf_issue = [&]() {   // one thread
        double *A = stage_base + pf_slot * SEP_STAGE_DOUBLES;
        constexpr uint32_t bytes = (SEP_A_TILE + SEP_B_TILE + (SCALED ? SEP_C_BLOCK : 0)) * sizeof(double);
        mbar_arrive_expect_tx(&full[pf_slot], bytes);
        bulk_g2s(A, pf_a + (int64_t)pf_s * SEP_A_TILE, SEP_A_TILE * sizeof(double), &full[pf_slot]);
        bulk_g2s(A + SEP_A_TILE, pf_b + (int64_t)pf_s * SEP_B_TILE, SEP_B_TILE * sizeof(double), &full[pf_slot]);
        if (SCALED)
            bulk_g2s(A + SEP_A_TILE + SEP_B_TILE, pf_c + (int64_t)pf_s * SEP_C_BLOCK,
                     SEP_C_BLOCK * sizeof(double), &full[pf_slot]);
    };
    auto pf_advance = [&]() {   // all threads, uniform
        if (++pf_slot == SEP_STAGES) { pf_slot = 0; ++pf_round; }
        if (++pf_s == n_stages) {
            pf_s = 0;
            pf_tile += gridDim.x;
            if (pf_tile < n_tiles) pf_decode();
        }
    };
    if (pf_tile < n_tiles) pf_decode();
#pragma unroll
    for (int p = 0; p < DEPTH; ++p) {
        if (pf_tile < n_tiles) {
            if (tid == 0) pf_issue();
            pf_advance();
        }
    }

    // warp tile 32 x 64 = 4 x 8 DMMA tiles of 8x8; fragment owner: g = lane/4, t = lane%4
    // Sub-partition k hosts warps k and k + 4: it gets BOTH 64-column halves of band k.  In a partial
    // column tile (last axis not a multiple of 128) the right half has fewer valid 8-column groups than the
    // left; with both on the same FP64 pipe the tile costs ceil(width / 8) / 16 of a full one.
    const int wr = warp & 3;                        // 0..3 : 32-row band
    const int wc = ((warp >> 2) ^ warp) & 1;        // 0..1 : 64-column half
    const int g = lane >> 2;
    const int t = lane & 3;
    const int a_off = (wr * 32 + g) * SEP_AST + t;               // + i*8*SEP_AST + 4*k4
    const int b_off = SEP_A_TILE + t * SEP_BST + wc * 64 + g;    // + 4*k4*SEP_BST + j*8

    int slot = 0;
    uint32_t round = 0;
    int turn = 0;                           // warp whose lane 0 issues the next prefetch
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        double acc[4][8][2];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

        // valid 8-column groups of this warp's half in this column tile (8 everywhere but in the last one)
        int jv = 8;
        if (PARTIAL) {
            const int cw_tile = (int)min((int64_t)SEP_TN, prm.lc - (int64_t)col_tile_of(tile) * SEP_TN);
            jv = max(0, min(8, (cw_tile - 64 * wc + 7) >> 3));
        }

        for (int s = 0; s < n_stages; ++s) {
            // prefetch DEPTH steps ahead into the slot of the step before last, which every warp
            // released long ago (the wait on its "empty" barrier practically never blocks)
            if (pf_tile < n_tiles) {
                if (lane == 0 && warp == turn) {
                    if (pf_round > 0) mbar_wait(&empty[pf_slot], (pf_round - 1) & 1);
                    pf_issue();
                }
                pf_advance();
            }
            turn = (turn + 1) & (SEP_WARPS - 1);
            __syncwarp();
            mbar_wait(&full[slot], round & 1);
            const double *S = stage_base + slot * SEP_STAGE_DOUBLES;
            auto contract_stage = [&](auto full_tag) {
                constexpr bool FULL = decltype(full_tag)::value;
#pragma unroll
                for (int k4 = 0; k4 < SEP_KC / 2; ++k4) {   // 4 contraction indices = 2 modes
                    double af[4], bf[8];
#pragma unroll
                    for (int i = 0; i < 4; ++i) af[i] = S[a_off + i * 8 * SEP_AST + 4 * k4];
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (FULL || j < jv) bf[j] = S[b_off + 4 * k4 * SEP_BST + j * 8];
                    if (SCALED) {
                        // this lane holds part (t & 1) of mode 2*k4 + (t >> 1): (cos, sin) of the last row
                        // axis.  With c = (cr, ci) the slow-axis factor, the contraction needs
                        //   part 0:  Re(c e) =  cr*cos - ci*sin       part 1: -Im(c e) = -cr*sin - ci*cos
                        // i.e. alpha*own + beta*partner with alpha = +-cr, beta = -ci.
                        const double2 cc = *reinterpret_cast<const double2 *>(
                            S + SEP_A_TILE + SEP_B_TILE + 2 * (2 * k4 + (t >> 1)));
                        const double alpha = (t & 1) ? -cc.x : cc.x;
                        const double beta = -cc.y;
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const double partner = __shfl_xor_sync(0xffffffffu, af[i], 1);
                            af[i] = fma(alpha, af[i], beta * partner);
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (FULL || j < jv) {
#pragma unroll
                            for (int i = 0; i < 4; ++i) dmma_884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
                        }
                }
            };
            if (!PARTIAL || jv == 8) contract_stage(std::true_type{});
            else if (jv > 0) contract_stage(std::false_type{});
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[slot]);
            if (++slot == SEP_STAGES) { slot = 0; ++round; }
        }

        // epilogue: thread holds C[g][2t], C[g][2t+1] of every 8x8 tile
        const int ct = col_tile_of(tile);
        const int64_t rest = tile / prm.n_col_tiles;
        const int rt = (int)(rest % prm.n_row_tiles);
        const int64_t z = rest / prm.n_row_tiles;
        const int comp = (int)(z % prm.ncomp);
        const int64_t fl = z / prm.ncomp;
        double *out = prm.out + ((prm.batch0 + fl) * prm.ncomp + comp) * prm.out_fstride;
        const bool vec2 = (prm.lc & 1) == 0 && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
        const int64_t col0 = (int64_t)ct * SEP_TN;
        int64_t row0, row_end;                      // first row of the tile, end of its valid rows
        if (SCALED) {
            const int64_t sl = (prm.rt0 + rt) / prm.n_ytiles;
            const int64_t iy0 = (int64_t)((prm.rt0 + rt) % prm.n_ytiles) * SEP_TM;
            row0 = sl * prm.ly + iy0;
            row_end = sl * prm.ly + prm.ly;         // rows of one slow index never spill into the next
        } else {
            row0 = prm.row_begin + (int64_t)rt * SEP_TM;
            row_end = prm.n_rows;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int64_t row = row0 + wr * 32 + i * 8 + g;
            if (row >= row_end) continue;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int64_t col = col0 + wc * 64 + j * 8 + 2 * t;
                double *dst = out + row * prm.lc + col;
                if (prm.epi.on) {
                    // (the per-point arrays are only read for columns inside the mesh)
                    const int64_t idx = row * prm.lc + col;
                    if (col < prm.lc) acc[i][j][0] = epi_apply(prm.epi, acc[i][j][0], comp, idx);
                    if (col + 1 < prm.lc) acc[i][j][1] = epi_apply(prm.epi, acc[i][j][1], comp, idx + 1);
                }
                if (vec2 && col + 1 < prm.lc) {
                    *reinterpret_cast<double2 *>(dst) = make_double2(acc[i][j][0], acc[i][j][1]);
                } else {
                    if (col < prm.lc) dst[0] = acc[i][j][0];
                    if (col + 1 < prm.lc) dst[1] = acc[i][j][1];
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
inline int launch_agen(const AgenParams &ap, int64_t n_batch_chunk, int sm_count, cudaStream_t st)
{
    const int n_stages = ap.n_modes_pad / SEP_KC;
    // enough CTAs to fill the machine a few times over, never more stage-splits than stages
    const int64_t ctas = (int64_t)ap.n_row_tiles * n_batch_chunk;
    const int ysplit = (int)std::min<int64_t>(n_stages, std::max<int64_t>(1, (8LL * sm_count + ctas - 1) / ctas));
    dim3 grid((unsigned)ap.n_row_tiles, (unsigned)ysplit, (unsigned)n_batch_chunk);
    // Same shared-memory carve-out as the contraction kernel, otherwise the two kernels cannot
    // be resident on one SM at the same time and the overlap is lost.
    static std::atomic<uint64_t> carveout_set{0};
    if (first_launch_on_device(carveout_set)) {
#define GSB_AGEN_ATTR(N) GSB_CUDA(cudaFuncSetAttribute(agen_kernel<N>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        GSB_AGEN_ATTR(1) GSB_AGEN_ATTR(2) GSB_AGEN_ATTR(3) GSB_AGEN_ATTR(4) GSB_AGEN_ATTR(5) GSB_AGEN_ATTR(6) GSB_AGEN_ATTR(7)
#undef GSB_AGEN_ATTR
    }
    switch (ap.n_row_axes) {
    case 1: agen_kernel<1><<<grid, SEP_TM, 0, st>>>(ap); break;
    case 2: agen_kernel<2><<<grid, SEP_TM, 0, st>>>(ap); break;
    case 3: agen_kernel<3><<<grid, SEP_TM, 0, st>>>(ap); break;
    case 4: agen_kernel<4><<<grid, SEP_TM, 0, st>>>(ap); break;
    case 5: agen_kernel<5><<<grid, SEP_TM, 0, st>>>(ap); break;
    case 6: agen_kernel<6><<<grid, SEP_TM, 0, st>>>(ap); break;
    case 7: agen_kernel<7><<<grid, SEP_TM, 0, st>>>(ap); break;
    default:
        return fail(GSB_ERR_ARGUMENT, "structured path needs 2 <= dim <= 8");
    }
    g_launches.fetch_add(1);
    GSB_CUDA(cudaGetLastError());
    return GSB_OK;
}

inline int launch_contract(ContractParams cp, int64_t n_batch_chunk, int sm_count, bool scaled, cudaStream_t st)
{
    cp.n_fields = n_batch_chunk * cp.ncomp;
    const int64_t n_tiles = (int64_t)cp.n_col_tiles * cp.n_row_tiles * cp.n_fields;
    dim3 grid((unsigned)std::min<int64_t>(n_tiles, sm_count));
    static std::atomic<uint64_t> attr_set{0};
    if (first_launch_on_device(attr_set)) {
        // full 228 KB carve-out: leaves room next to this CTA for A-generation CTAs
#define GSB_CONTRACT_ATTR(K)                                                                                         \
        GSB_CUDA(cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SEP_SMEM_BYTES));           \
        GSB_CUDA(cudaFuncSetAttribute(K, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        GSB_CONTRACT_ATTR((contract_kernel<false, false>)) GSB_CONTRACT_ATTR((contract_kernel<false, true>))
        GSB_CONTRACT_ATTR((contract_kernel<true, false>)) GSB_CONTRACT_ATTR((contract_kernel<true, true>))
#undef GSB_CONTRACT_ATTR
    }
    const bool partial = (cp.lc % SEP_TN) != 0 && !cp.no_partial;
    if (scaled) {
        if (partial) contract_kernel<true, true><<<grid, SEP_THREADS, SEP_SMEM_BYTES, st>>>(cp);
        else contract_kernel<true, false><<<grid, SEP_THREADS, SEP_SMEM_BYTES, st>>>(cp);
    } else {
        if (partial) contract_kernel<false, true><<<grid, SEP_THREADS, SEP_SMEM_BYTES, st>>>(cp);
        else contract_kernel<false, false><<<grid, SEP_THREADS, SEP_SMEM_BYTES, st>>>(cp);
    }
    g_launches.fetch_add(1);
    GSB_CUDA(cudaGetLastError());
    return GSB_OK;
}

// device-side mesh expansion for meshes too small for the tiled kernel:
// pos[t][r] = sum_s M[t][s] a_s[i_s(r)], r in C order (generate_grid + isometrize).
struct ExpandParams {
    const double *axes;
    int64_t axis_off[GSB_MAX_DIM];
    int64_t axis_len[GSB_MAX_DIM];
    double matrix[GSB_MAX_DIM * GSB_MAX_DIM];
    int dim;
    int64_t n;
    double *pos;  // (dim, n)
};

__global__ void expand_grid_kernel(const ExpandParams ep)
{
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < ep.n;
         r += (int64_t)gridDim.x * blockDim.x) {
        double g[GSB_MAX_DIM];
        int64_t rem = r;
        for (int t = ep.dim - 1; t >= 0; --t) {
            const int64_t it = rem % ep.axis_len[t];
            rem /= ep.axis_len[t];
            g[t] = ep.axes[ep.axis_off[t] + it];
        }
        for (int t = 0; t < ep.dim; ++t) {
            // same left-to-right order as np.dot(matrix, pos) row t
            double v = 0.0;
            for (int s = 0; s < ep.dim; ++s) v += ep.matrix[t * ep.dim + s] * g[s];
            ep.pos[(int64_t)t * ep.n + r] = v;
        }
    }
}

}  // namespace gsb
