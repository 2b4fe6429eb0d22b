// gsb_sepk.cuh -- separable summation on structured meshes, second generation ("stream-K" contraction).
//
// Same mathematics as gsb_separable.cuh (reference: src/gstools/field/generator.py:193-199 on the mesh of
// src/gstools/tools/geometric.py:340-356): with k' = M^T k the phase splits per axis and
//     u[row, c] = sum_j Re( A_j(row) E_j(c) ),   A_j(row) = c_j(slow) T_j(iy),   E_j(c) = exp(i k'_last a_last[c])
// where the row index is split as row = slow * ly + iy:
//     T_j(iy)   = exp(i sum_{t in tile axes} k'_t a_t[i_t])   the TILE-AXIS table: the trailing row axes folded into
//                                                              one axis of ly entries (pre-tiled, L2 resident)
//     c_j(slow) = (z1_j - i z2_j) sf_j prod_{t in prefix axes} exp(i k'_t a_t[i_t])   one complex factor per
//                                                              (slow index, mode): 128 bytes per pipeline stage
// What changed against the first generation, and why:
//   * NO pre-generated A operand.  The first generation wrote A = c T to HBM (5.2 GB for the 512^3 field) from a
//     second kernel and re-read it; here every warp rescales the T fragments of ITS OWN rows right after
//     loading them.  Two layout changes make that cheap: (1) a warp owns 16 rows x 128 columns (2 x 16 DMMA
//     tiles) instead of 32 x 64, so no two warps rescale the same rows -- half the FP64 work; (2) within a
//     stage the contraction index is ordered [4 modes: real parts][the same 4 modes: imaginary parts], so the lane
//     that feeds k-slot t of the DMMA holds BOTH cos and sin of mode t (one LDS.128) and needs no shuffle.
//     Cost: 2 DMUL + 2 DFMA per 32 DMMA.8x8x4 per warp (1.6 % of the FP64 issue slots).
//   * STREAM-K.  The first generation walked whole 128x128 output tiles (tile t, t + grid, ...): a 128^3 field is
//     128 tiles on 148 SMs (one partial wave), 200^3 pads 200 -> 256 columns, and every row chunk was a separate
//     launch with its own pipeline fill and drain.  Here ONE persistent launch splits the (tile, stage) iteration
//     space -- weighted by what a tile really costs when it hangs over the mesh edge -- into equal contiguous
//     shares, one per SM.  A tile that straddles two shares is finished by the CTA that started it: the later
//     CTA(s) write their partial accumulators to a scratch slot and raise a flag, the owner adds them in a fixed
//     order (deterministic, no atomics on data).
//   * ROW PACKING (round 2, second step).  Row tiles used to restart at every slow index: a 200^3 mesh computed
//     256 rows per plane for 200 (78 %), 300^3 384 for 300.  Now the rows of a field live in ONE virtual row space,
//     `lyp` rows per slow index (ly rounded up to 8, the height of a DMMA row group; or to 128 when nothing is
//     packed), cut into 128-row tiles regardless of the slow boundaries.  A tile that spans several slow indices
//     gets its T rows as several bulk copies (one per slow index, the rows are contiguous in the table) and one
//     128-byte block of slow-axis factors per slow index; every 8-row group rescales with the factors of ITS slow
//     index.  Costs one more LDS.128 per quad of modes; saves the padding rows.
#pragma once

#include <algorithm>
#include <type_traits>
#include <vector>

#include "gsb_common.cuh"

namespace gsb {

constexpr int SK_TM = 128;                 // rows per tile
constexpr int SK_TN = 128;                 // columns per tile
constexpr int SK_KC = 8;                   // modes per pipeline stage (multiple of 4)
constexpr int SK_STAGES = 4;
constexpr int SK_WARPS = 8;
constexpr int SK_THREADS = SK_WARPS * 32;
static_assert(SK_KC % 4 == 0, "a stage holds whole quads of modes");
// shared-memory / global tile layouts (doubles)
//   T tile [128 rows][SK_AST]: (cos, sin) -- or (Re A, -Im A) when nothing is rescaled -- of mode kc at 2*kc;
//                              row stride 24 makes the LDS.128 of a quarter warp hit 8 distinct 16-byte banks
//   B tile [2*KC rows][SK_BST]: row 8*(kc/4) + 4*part + kc%4 (part 0: p cos, part 1: p sin), stride 132
//   C block [KC] complex slow-axis factors
constexpr int SK_AST = 2 * SK_KC + 8;
constexpr int SK_BST = SK_TN + 4;
constexpr int SK_A_TILE = SK_TM * SK_AST;
constexpr int SK_B_TILE = 2 * SK_KC * SK_BST;
constexpr int SK_C_BLOCK = 2 * SK_KC;
constexpr int SK_MAXSEG = 4;               // slow indices a packed row tile may touch (rows per slow index >= 48)
constexpr int SK_PACK_MIN_ROWS = 48;
constexpr int SK_STAGE_DOUBLES = SK_A_TILE + SK_B_TILE + SK_MAXSEG * SK_C_BLOCK;
constexpr size_t SK_SMEM_BYTES =
    (size_t)SK_STAGES * SK_STAGE_DOUBLES * sizeof(double) + 2 * SK_STAGES * sizeof(uint64_t) + 128;
constexpr int SK_MAX_AXES = GSB_MAX_DIM;
constexpr int SK_MAX_GRID = 160;           // CTAs of the persistent grid (one per SM; B200: 148)

// ---------------------------------------------------------------------------------------------
// tables
// ---------------------------------------------------------------------------------------------
struct SkTableParams {
    const double *cov, *z1, *z2, *sf;   // (n_batch, dim, n_modes), (n_batch, n_modes) x 3 (sf optional)
    const double *axes;                 // concatenated axis coordinates
    int64_t axis_off[SK_MAX_AXES];      // per MESH axis
    int64_t axis_len[SK_MAX_AXES];
    double matrix[GSB_MAX_DIM * GSB_MAX_DIM];
    int dim;                            // mesh axes = dimension of the wave vectors
    // mesh axes, in order: [outer slow axes | tile axes | inner slow axes | column axes]
    int n_prefix;                       // outer slow axes
    int n_tile_axes;                    // folded into the tile axis (ly entries)
    int n_inner;                        // inner slow axes (n_in entries); slow = outer * n_in + inner
    int n_col_axes;                     // trailing axes (1 or 2) folded into the column axis (lc entries)
    int64_t ly, lc, n_slow, n_in;
    int64_t n_modes;
    int n_modes_pad, ncomp;
    int n_ytiles, n_col_tiles;
    double *ttab;                       // [b][stage][n_ytiles * 128 rows][SK_AST]: all rows of a stage are contiguous
    double *btile;                      // [b][comp][col tile][stage] blocks of SK_B_TILE
    double2 *ctab;                      // [b][slow][mode] slow-axis factors; NULL: no prefix axes, the T table
                                        //   carries (z1 - i z2) sf itself and stores (Re A, -Im A)
    unsigned *flags;                    // stream-K flags, zeroed here (n_flags entries)
    int n_flags;
    int small_index;                    // every table of the call has < 2^31 entries: 32-bit entry arithmetic
};

__device__ __forceinline__ double sk_kprime(const SkTableParams &tp, const double *cov, int t, int64_t j)
{
    // k'_t = sum_s M[s][t] k_s   (phase = k . (M a) = (M^T k) . a)
    double kp = 0.0;
    for (int s = 0; s < tp.dim; ++s) kp = fma(tp.matrix[s * tp.dim + t], cov[(int64_t)s * tp.n_modes + j], kp);
    return kp;
}

// phase of mode j at entry i of the axis obtained by folding mesh axes [t0, t0 + nt) (last fastest)
template <typename I>
__device__ __forceinline__ double sk_folded_phase(const SkTableParams &tp, const double *cov, int t0, int nt, I i, I j)
{
    double phase = 0.0;
    I rem = i;
    for (int t = t0 + nt - 1; t >= t0; --t) {
        const I len = (I)tp.axis_len[t];
        const I it = rem % len;
        rem /= len;
        phase = fma(sk_kprime(tp, cov, t, (int64_t)j), tp.axes[tp.axis_off[t] + (int64_t)it], phase);
    }
    return phase;
}

// All tables of a call in one launch.  blockIdx.y: 0 = T table (tile axis), 1 = B table (column axis),
// 2 = slow-axis factors;  blockIdx.z = batch entry.  Full-accuracy sincos (libdevice), O((ly + lc + n_slow) N).
// The padding columns of the pre-tiled blocks travel with the bulk copies: they are written (zero) here too.
// I: index type of the entry loop -- 32-bit whenever every table has fewer than 2^31 entries (always, in practice):
// the two divisions per entry that decode (axis entry, mode) and those of folded axes then cost a quarter of their
// 64-bit versions, which is a third of this kernel's time.
template <typename I>
__device__ __forceinline__ void sk_tables_body(const SkTableParams &tp)
{
    const int which = blockIdx.y;
    const int64_t b = blockIdx.z;
    const double *cov = tp.cov + b * tp.dim * tp.n_modes;
    const int n_stages = tp.n_modes_pad / SK_KC;
    if (which == 0 && b == 0 && blockIdx.x == 0)
        for (int i = threadIdx.x; i < tp.n_flags; i += blockDim.x) tp.flags[i] = 0u;
    I width;
    if (which == 0) width = (I)((int64_t)tp.n_ytiles * SK_TM);
    else if (which == 1) width = (I)((int64_t)tp.n_col_tiles * SK_TN);
    else width = (I)tp.n_slow;
    const I n_pad = (I)tp.n_modes_pad;
    const I total = width * n_pad;
    for (I idx = (I)(blockIdx.x * (int64_t)blockDim.x + threadIdx.x); idx < total; idx += (I)((int64_t)gridDim.x * blockDim.x)) {
        const bool mode_major = which != 2;     // T / B: consecutive threads walk along the axis; C: along the modes
        const I j = mode_major ? idx / width : idx % n_pad;
        const I i = mode_major ? idx - j * width : idx / n_pad;
        const bool live = (int64_t)j < tp.n_modes;
        if (which == 0) {
            double c = 0.0, s = 0.0;
            if (live && (int64_t)i < tp.ly) {
                sincos(sk_folded_phase<I>(tp, cov, tp.n_prefix, tp.n_tile_axes, i, j), &s, &c);
                if (!tp.ctab) {
                    const double w = tp.sf ? tp.sf[b * tp.n_modes + (int64_t)j] : 1.0;
                    const double a = w * tp.z1[b * tp.n_modes + (int64_t)j], bb = w * tp.z2[b * tp.n_modes + (int64_t)j];
                    const double re = a * c + bb * s;      // (a - i bb)(c + i s)
                    const double im = a * s - bb * c;
                    c = re;
                    s = -im;
                }
            }
            const int yt = (int)(i / SK_TM), r = (int)(i % SK_TM);
            const int kc = (int)(j % SK_KC);
            double *row = tp.ttab + ((b * n_stages + (int64_t)(j / SK_KC)) * tp.n_ytiles + yt) * (int64_t)SK_A_TILE + r * SK_AST;
            *reinterpret_cast<double2 *>(row + 2 * kc) = make_double2(c, s);
            if (kc == 0) {
#pragma unroll
                for (int u = 2 * SK_KC; u < SK_AST; u += 2)
                    *reinterpret_cast<double2 *>(row + u) = make_double2(0.0, 0.0);
            }
        } else if (which == 1) {
            double c = 0.0, s = 0.0;
            if (live && (int64_t)i < tp.lc)
                sincos(sk_folded_phase<I>(tp, cov, tp.n_prefix + tp.n_tile_axes + tp.n_inner, tp.n_col_axes, i, j), &s, &c);
            double k2 = 0.0, k0 = 0.0;
            if (tp.ncomp > 1 && live) {
                for (int s2 = 0; s2 < tp.dim; ++s2) {
                    const double k = cov[(int64_t)s2 * tp.n_modes + (int64_t)j];
                    k2 += k * k;
                }
                k0 = cov[(int64_t)j];
            }
            const int ct = (int)(i / SK_TN), col = (int)(i % SK_TN);
            const int st = (int)(j / SK_KC), kc = (int)(j % SK_KC);
            const int row = 8 * (kc / 4) + (kc % 4);
            for (int comp = 0; comp < tp.ncomp; ++comp) {
                double p = 1.0;
                if (tp.ncomp > 1) {
                    // incompressible projector on the ORIGINAL wave vector (generator.py:479-495)
                    p = 0.0;
                    if (live) p = ((comp == 0) ? 1.0 : 0.0) - cov[(int64_t)comp * tp.n_modes + (int64_t)j] * k0 / k2;
                }
                double *tile = tp.btile +
                               (((b * tp.ncomp + comp) * tp.n_col_tiles + ct) * n_stages + st) * (int64_t)SK_B_TILE;
                tile[row * SK_BST + col] = p * c;
                tile[(row + 4) * SK_BST + col] = p * s;
                if (col < SK_BST - SK_TN) {
                    tile[row * SK_BST + SK_TN + col] = 0.0;
                    tile[(row + 4) * SK_BST + SK_TN + col] = 0.0;
                }
            }
        } else {
            // c[b][slow][j] = (z1_j - i z2_j) sf_j exp(i sum_{slow axes} k'_t a_t[i_t(slow)])
            double2 e = make_double2(0.0, 0.0);
            if (live) {
                double c, s;
                const I n_in = (I)tp.n_in;
                const double phase = sk_folded_phase<I>(tp, cov, 0, tp.n_prefix, i / n_in, j) +
                                     sk_folded_phase<I>(tp, cov, tp.n_prefix + tp.n_tile_axes, tp.n_inner, i % n_in, j);
                sincos(phase, &s, &c);
                const double w = tp.sf ? tp.sf[b * tp.n_modes + (int64_t)j] : 1.0;
                const double a = w * tp.z1[b * tp.n_modes + (int64_t)j], bb = w * tp.z2[b * tp.n_modes + (int64_t)j];
                e = make_double2(a * c + bb * s, a * s - bb * c);
            }
            tp.ctab[(b * tp.n_slow + (int64_t)i) * tp.n_modes_pad + (int64_t)j] = e;
        }
    }
}

__global__ void sk_tables_kernel(const SkTableParams tp)
{
    if (tp.small_index) sk_tables_body<unsigned>(tp);
    else sk_tables_body<int64_t>(tp);
}

// ---------------------------------------------------------------------------------------------
// stream-K plan (host).  Tiles are numbered  tile = ((z * n_slow + slow) * n_ytiles + yt) * n_col_tiles + ct
// (z = field * ncomp + comp).  A pipeline stage of a tile costs rowq(yt) * colg(ct) units: rowq in 1..4 counts
// the passes of 8-row groups each of the 4 FP64 pipes makes (row groups are dealt round-robin to the pipes),
// colg in {2, 4, ..., 16} the 8-column groups computed (inside the mesh, rounded up to 2); floor: sk_tile_cost.  Boundaries of the equal-cost shares are snapped to stages.
// ---------------------------------------------------------------------------------------------
struct SkBound { int64_t tile; int32_t stage; int32_t pad; };

inline int sk_rowq(int64_t ly, int yt)
{
    const int64_t rows = std::min<int64_t>(SK_TM, ly - (int64_t)yt * SK_TM);
    const int groups = (int)((rows + 7) / 8);
    return (groups + 3) / 4;
}
inline int sk_colg(int64_t lc, int ct)
{
    const int64_t cols = std::min<int64_t>(SK_TN, lc - (int64_t)ct * SK_TN);
    return 2 * (int)((cols + 15) / 16);      // the kernel has compile-time variants for 2, 4, ..., 16 column groups
}
// Cost of one pipeline stage of tile (yt, ct).  However little of a tile lies inside the mesh, its stage still
// moves the whole 41.6 KB of operands from L2 to shared memory: measured (profiles/r02_ncu_sk_*): half-height
// tiles pull 4.85 TB/s through the L2 -> SM path and run at 84 % of the DMMA rate, full tiles need 2.7 TB/s.
// The floor of 40 units (a full tile stage is 64) stands for that bandwidth limit.
constexpr int SK_COST_FLOOR = 40;
inline int sk_tile_cost(int64_t ly, int64_t lc, int yt, int ct)
{
    return std::max(SK_COST_FLOOR, sk_rowq(ly, yt) * sk_colg(lc, ct));
}

// Shares of the tiles [tile_begin, tile_end) for `grid` CTAs: bnd[0..grid], bnd[c] <= bnd[c+1],
// bnd[0] = (tile_begin, 0), bnd[grid] = (tile_end, 0).  Returns the grid size actually used (<= max_grid).
inline int sk_plan(int64_t tile_begin, int64_t tile_end, int n_ytiles, int n_col_tiles, int n_stages, int64_t ly,
                   int64_t lc, int max_grid, std::vector<SkBound> &bnd)
{
    max_grid = std::max(1, std::min(max_grid, SK_MAX_GRID));
    const int period = n_ytiles * n_col_tiles;
    std::vector<int64_t> prefix((size_t)period + 1, 0);     // cost of the first p tiles of a period
    std::vector<int> w((size_t)period);
    for (int p = 0; p < period; ++p) {
        w[p] = sk_tile_cost(ly, lc, p / n_col_tiles, p % n_col_tiles);
        prefix[p + 1] = prefix[p] + (int64_t)w[p] * n_stages;
    }
    const int64_t P = prefix[period];
    auto cum = [&](int64_t tile) { return (tile / period) * P + prefix[(size_t)(tile % period)]; };
    const int64_t c0 = cum(tile_begin), c1 = cum(tile_end);
    // A share is never smaller than an eighth of a full tile: the owner of a tile adds the partial accumulators
    // of the other contributors one after the other (128 KB each), so a tiny mesh (100 x 100: ONE tile) must not
    // be cut into 125 single-stage shares.
    const int64_t min_share = std::max<int64_t>(64, (int64_t)n_stages * 64 / 8);
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(max_grid, (c1 - c0) / min_share));
    bnd.assign((size_t)grid + 1, SkBound{tile_end, 0, 0});
    bnd[0] = SkBound{tile_begin, 0, 0};
    for (int c = 1; c < grid; ++c) {
        // target cost from the start of the numbering; (c1 - c0) * c fits: cost < 2^50, c < 2^10
        const int64_t target = c0 + (c1 - c0) / grid * c + (c1 - c0) % grid * c / grid;
        const int64_t per = target / P, rem = target % P;
        const int p = (int)(std::upper_bound(prefix.begin(), prefix.end(), rem) - prefix.begin()) - 1;   // prefix[p] <= rem
        const int64_t tile = per * period + p;
        const int stage = (int)((rem - prefix[(size_t)p]) / w[(size_t)p]);
        bnd[(size_t)c] = SkBound{tile, stage, 0};
        if (bnd[(size_t)c].tile >= tile_end) bnd[(size_t)c] = SkBound{tile_end, 0, 0};
    }
    return grid;
}

// ---------------------------------------------------------------------------------------------
// the contraction
// ---------------------------------------------------------------------------------------------
struct SkParams {
    const double *ttab, *btile;
    const double2 *ctab;
    int n_ytiles, n_col_tiles, n_stages, ncomp;
    int64_t n_slow, n_in;     // slow = outer * n_in + inner: row of (slow, iy) = (outer * ly + iy) * n_in + inner
    int64_t ly, lc;
    // virtual row space of a field: slow index s owns the rows [s * lyp, s * lyp + ly); tile vt covers the virtual
    // rows [128 vt, 128 vt + 128).  lyp = n_ytiles * 128 (nothing packed) or ly rounded up to 8 (packed).
    int64_t lyp, vrows;       // vrows = n_slow * lyp
    int64_t n_vt;             // row tiles per field = ceil(vrows / 128)
    int n_modes_pad;
    double *out;              // field (batch, comp) at out + (batch*ncomp + comp)*out_fstride; row r at + r*lc
    int64_t out_fstride;
    SkBound bnd[SK_MAX_GRID + 1];   // shares of the gridDim.x CTAs (kernel parameter: constant bank, no upload)
    double *slots;            // [gridDim.x][SK_TM*SK_TN] partial accumulators of a CTA's leading (non-owned) segment
    unsigned *flags;          // [gridDim.x], zeroed before the launch
    Epi epi;
};

__device__ __forceinline__ void sk_dmma(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ unsigned sk_ld_acquire(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void sk_st_release(unsigned *p, unsigned v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// One pipeline stage of a warp: NI row groups (1 or 2) x JV column groups (4, 8, 12 or 16), all compile-time, so
// that partial tiles run the same straight-line, software-pipelined code as full ones (a first version predicated
// every column group at run time: each LDS was followed by its dependent DMMA, 2x slower on partial tiles).
// PACK: the two row groups of the warp may belong to different slow indices -> two blocks of slow-axis factors.
template <bool SCALE, bool PACK, int JV, int NI>
__device__ __forceinline__ void sk_stage(const double *S, double (&acc)[2][16][2], int a_off0, int a_off1, int b_off,
                                         int c_off0, int c_off1)
{
#pragma unroll
    for (int q = 0; q < SK_KC / 4; ++q) {
        // (cos, sin) of mode 4q + t for this lane's rows; rescaled by the slow-axis factor
        double ar[2], ai[2];
        const double2 e0 = *reinterpret_cast<const double2 *>(S + a_off0 + 8 * q);
        double2 e1 = make_double2(0.0, 0.0);
        if (NI > 1) e1 = *reinterpret_cast<const double2 *>(S + a_off1 + 8 * q);
        if (SCALE) {
            const double2 cc = *reinterpret_cast<const double2 *>(S + c_off0 + 8 * q);
            //  Re(c e) = cr cos - ci sin        -Im(c e) = -(cr sin + ci cos)
            ar[0] = fma(cc.x, e0.x, -(cc.y * e0.y));
            ai[0] = fma(-cc.x, e0.y, -(cc.y * e0.x));
            if (NI > 1) {
                double2 c1 = cc;
                if (PACK) c1 = *reinterpret_cast<const double2 *>(S + c_off1 + 8 * q);
                ar[1] = fma(c1.x, e1.x, -(c1.y * e1.y));
                ai[1] = fma(-c1.x, e1.y, -(c1.y * e1.x));
            }
        } else {
            ar[0] = e0.x; ai[0] = e0.y;
            ar[1] = e1.x; ai[1] = e1.y;
        }
        const double *Bq = S + b_off + 8 * q * SK_BST;
#pragma unroll
        for (int part = 0; part < 2; ++part) {
#pragma unroll
            for (int j = 0; j < JV; ++j) {
                const double bf = Bq[4 * part * SK_BST + 8 * j];
                sk_dmma(acc[0][j][0], acc[0][j][1], part ? ai[0] : ar[0], bf);
                if (NI > 1) sk_dmma(acc[1][j][0], acc[1][j][1], part ? ai[1] : ar[1], bf);
            }
        }
    }
}

// Row geometry of tile vt (warp-uniform): the segments of the virtual row space it covers.  Segment k holds
// `rows[k]` tile rows from tile row `row0[k]` on, they are the rows iy0[k] .. of slow index slow[k].
template <bool PACK>
struct SkTileGeo {
    static constexpr int NS = PACK ? SK_MAXSEG : 1;
    int nseg;
    int groups;          // 8-row groups of the tile that hold mesh rows (a prefix of the 16)
    int row0[NS], rows[NS], iy0[NS];
    int64_t slow[NS];
};

template <bool PACK>
__device__ __forceinline__ void sk_tile_geo(const SkParams &prm, int64_t vt, SkTileGeo<PACK> &g)
{
    const int64_t v0 = vt * SK_TM;
    int64_t slow = v0 / prm.lyp;
    int iy = (int)(v0 - slow * prm.lyp);
    if (!PACK) {
        // lyp is a multiple of 128: the tile lies inside one slow index; rows beyond ly are padding
        const int rows = (int)min((int64_t)SK_TM, prm.ly - iy);
        g.nseg = 1;
        g.groups = (rows + 7) >> 3;
        g.row0[0] = 0;
        g.rows[0] = g.groups * 8;
        g.iy0[0] = iy;
        g.slow[0] = slow;
        return;
    }
    // lyp = ly rounded up to 8: every group of a slow index holds at least one mesh row
    const int avail = (int)min((int64_t)SK_TM, prm.vrows - v0);
    int covered = 0, n = 0;
#pragma unroll
    for (int k = 0; k < SK_MAXSEG; ++k) {
        if (covered < avail) {
            const int take = min(avail - covered, (int)prm.lyp - iy);
            g.row0[k] = covered;
            g.rows[k] = take;
            g.iy0[k] = iy;
            g.slow[k] = slow;
            covered += take;
            n = k + 1;
            ++slow;
            iy = 0;
        }
    }
    g.nseg = n;
    g.groups = avail >> 3;
}

// SCALE   : rescale the T fragments by the slow-axis factor (meshes with prefix axes); false: the T table is A itself
// PARTIAL : tiles may hang over the mesh edge: skip the row / column groups outside
// PACK    : row tiles span slow indices (see the file header); implies SCALE and PARTIAL
template <bool SCALE, bool PARTIAL, bool PACK>
__global__ void __launch_bounds__(SK_THREADS, 1) sk_contract_kernel(const __grid_constant__ SkParams prm)
{
    static_assert(!PACK || (SCALE && PARTIAL), "packing needs slow axes and handles partial tiles");
    extern __shared__ __align__(128) unsigned char sk_smem_raw[];
    double *stage_base = reinterpret_cast<double *>(sk_smem_raw);
    uint64_t *full = reinterpret_cast<uint64_t *>(stage_base + SK_STAGES * SK_STAGE_DOUBLES);
    uint64_t *empty = full + SK_STAGES;

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    const int n_stages = prm.n_stages;
    const int cta = blockIdx.x;
    const int64_t lyt = (int64_t)prm.n_ytiles * SK_TM;           // table rows per stage

    if (tid == 0) {
        for (int s = 0; s < SK_STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], SK_WARPS);
        }
        fence_barrier_init();
    }
    __syncthreads();

    // this CTA's share: (tile, stage) from lo (inclusive) to hi (exclusive)
    const SkBound lo = prm.bnd[cta], hi = prm.bnd[cta + 1];
    const int64_t last_tile = (hi.stage > 0) ? hi.tile : hi.tile - 1;      // last tile touched
    const bool has_work = (lo.tile < hi.tile) || (lo.tile == hi.tile && lo.stage < hi.stage);

    // ---- prefetch cursor: walks the same (tile, stage) sequence DEPTH steps ahead ----
    constexpr int DEPTH = SK_STAGES - 2;
    int64_t pf_tile = lo.tile;
    int pf_s = lo.stage, pf_send = 0;
    bool pf_live = has_work;
    int pf_slot = 0;
    uint32_t pf_round = 0;
    const double *pf_a = nullptr, *pf_b = nullptr, *pf_c = nullptr;
    int64_t pf_vt = 0;
    int pf_arows = SK_TM;
    auto pf_decode = [&]() {
        const int ct = (int)(pf_tile % prm.n_col_tiles);
        const int64_t rest = pf_tile / prm.n_col_tiles;
        pf_vt = rest % prm.n_vt;
        const int64_t z = rest / prm.n_vt;
        const int comp = (int)(z % prm.ncomp);
        const int64_t b = z / prm.ncomp;
        pf_a = prm.ttab + b * (int64_t)n_stages * lyt * SK_AST;
        pf_b = prm.btile + (((b * prm.ncomp + comp) * prm.n_col_tiles + ct) * (int64_t)n_stages) * SK_B_TILE;
        if (SCALE) pf_c = reinterpret_cast<const double *>(prm.ctab + b * prm.n_slow * prm.n_modes_pad);
        if (!PACK) {
            SkTileGeo<false> geo;
            sk_tile_geo<false>(prm, pf_vt, geo);
            pf_a += (int64_t)geo.iy0[0] * SK_AST;
            if (SCALE) pf_c += 2 * geo.slow[0] * prm.n_modes_pad;
            pf_arows = PARTIAL ? geo.rows[0] : SK_TM;
        }
        pf_send = (pf_tile == hi.tile) ? hi.stage : n_stages;
    };
    auto pf_issue = [&]() {   // one thread
        double *A = stage_base + pf_slot * SK_STAGE_DOUBLES;
        const double *ta = pf_a + (int64_t)pf_s * lyt * SK_AST;
        if (!PACK) {
            const uint32_t a_bytes = (uint32_t)(pf_arows * SK_AST * sizeof(double));
            const uint32_t bytes = a_bytes + (SK_B_TILE + (SCALE ? SK_C_BLOCK : 0)) * (uint32_t)sizeof(double);
            mbar_arrive_expect_tx(&full[pf_slot], bytes);
            bulk_g2s(A, ta, a_bytes, &full[pf_slot]);
            bulk_g2s(A + SK_A_TILE, pf_b + (int64_t)pf_s * SK_B_TILE, SK_B_TILE * sizeof(double), &full[pf_slot]);
            if (SCALE)
                bulk_g2s(A + SK_A_TILE + SK_B_TILE, pf_c + (int64_t)pf_s * SK_C_BLOCK, SK_C_BLOCK * sizeof(double),
                         &full[pf_slot]);
        } else {
            // the geometry is recomputed by the issuing lane (two integer divisions once per stage and CTA) instead of
            // living in registers of all 256 threads across the stage loop
            SkTileGeo<true> geo;
            sk_tile_geo<true>(prm, pf_vt, geo);
            const uint32_t bytes = (uint32_t)((geo.groups * 8 * SK_AST + SK_B_TILE + geo.nseg * SK_C_BLOCK) * sizeof(double));
            mbar_arrive_expect_tx(&full[pf_slot], bytes);
#pragma unroll
            for (int k = 0; k < SK_MAXSEG; ++k) {
                if (k < geo.nseg) {
                    bulk_g2s(A + geo.row0[k] * SK_AST, ta + (int64_t)geo.iy0[k] * SK_AST,
                             (uint32_t)(geo.rows[k] * SK_AST * sizeof(double)), &full[pf_slot]);
                    bulk_g2s(A + SK_A_TILE + SK_B_TILE + k * SK_C_BLOCK,
                             pf_c + 2 * geo.slow[k] * prm.n_modes_pad + (int64_t)pf_s * SK_C_BLOCK,
                             SK_C_BLOCK * sizeof(double), &full[pf_slot]);
                }
            }
            bulk_g2s(A + SK_A_TILE, pf_b + (int64_t)pf_s * SK_B_TILE, SK_B_TILE * sizeof(double), &full[pf_slot]);
        }
    };
    auto pf_advance = [&]() {   // all threads, uniform
        if (++pf_slot == SK_STAGES) { pf_slot = 0; ++pf_round; }
        if (++pf_s == pf_send) {
            pf_s = 0;
            ++pf_tile;
            if (pf_tile > last_tile) pf_live = false;
            else pf_decode();
        }
    };
    if (pf_live) pf_decode();
#pragma unroll
    for (int p = 0; p < DEPTH; ++p) {
        if (pf_live) {
            if (tid == 0) pf_issue();
            pf_advance();
        }
    }

    // fragment owner: g = lane / 4 (row within an 8-row group / column within an 8-column group), t = lane % 4
    // (k slot).  Warp w owns row groups w and w + 8: rows 8w .. 8w+7 and 64+8w .. 64+8w+7, all 128 columns.
    // Sub-partition s hosts warps s and s + 4, i.e. row groups s, s+4, s+8, s+12: the row groups of a tile that
    // hangs over the mesh edge are spread evenly over the four FP64 pipes.
    const int g = lane >> 2;
    const int t = lane & 3;
    const int a_off0 = (8 * warp + g) * SK_AST + 2 * t;            // + q * 8   (q = quad of modes)
    const int a_off1 = (8 * (warp + 8) + g) * SK_AST + 2 * t;
    const int b_off = SK_A_TILE + t * SK_BST + g;                  // + (8q + 4 part) * SK_BST + 8 j
    const int c_base = SK_A_TILE + SK_B_TILE + 2 * t;              // + segment * SK_C_BLOCK + q * 8

    int slot = 0;
    uint32_t round = 0;
    int turn = 0;
    for (int64_t tile = lo.tile; has_work && tile <= last_tile; ++tile) {
        const int s_begin = (tile == lo.tile) ? lo.stage : 0;
        const int s_end = (tile == hi.tile) ? hi.stage : n_stages;
        const int ct = (int)(tile % prm.n_col_tiles);
        const int64_t rest = tile / prm.n_col_tiles;
        const int64_t vt = rest % prm.n_vt;
        const int64_t z = rest / prm.n_vt;

        double acc[2][16][2];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

        // compile-time variants of the stage: column groups rounded up to 2 (the B tile is zero beyond the mesh),
        // row groups of this warp inside the mesh; variant = 2 * (jv2 - 1) + (ni - 1), 15 = the full tile, -1 = idle
        int variant = 15;
        int c_off0 = c_base, c_off1 = c_base;
        if (PARTIAL) {
            SkTileGeo<PACK> geo;
            sk_tile_geo<PACK>(prm, vt, geo);
            const int cols = (int)min((int64_t)SK_TN, prm.lc - (int64_t)ct * SK_TN);
            const int jv2 = (cols + 15) >> 4;
            const int ni = (warp < geo.groups ? 1 : 0) + (warp + 8 < geo.groups ? 1 : 0);
            variant = ni > 0 ? 2 * (jv2 - 1) + (ni - 1) : -1;
            if (PACK) {
                // the segment (= block of slow-axis factors) of each of the warp's row groups
                int s0 = 0, s1 = 0;
#pragma unroll
                for (int k = 1; k < SK_MAXSEG; ++k) {
                    if (k < geo.nseg && geo.row0[k] <= 8 * warp) s0 = k;
                    if (k < geo.nseg && geo.row0[k] <= 8 * (warp + 8)) s1 = k;
                }
                c_off0 = c_base + s0 * SK_C_BLOCK;
                c_off1 = c_base + s1 * SK_C_BLOCK;
            }
        }

        for (int s = s_begin; s < s_end; ++s) {
            if (pf_live) {
                if (lane == 0 && warp == turn) {
                    if (pf_round > 0) mbar_wait(&empty[pf_slot], (pf_round - 1) & 1);
                    pf_issue();
                }
                pf_advance();
            }
            turn = (turn + 1) & (SK_WARPS - 1);
            __syncwarp();
            mbar_wait(&full[slot], round & 1);
            const double *S = stage_base + slot * SK_STAGE_DOUBLES;
            if (!PARTIAL || variant == 15) {
                sk_stage<SCALE, PACK, 16, 2>(S, acc, a_off0, a_off1, b_off, c_off0, c_off1);
            } else {
#define GSB_SK_CASE(JV)                                                                                               \
                case 2 * (JV / 2 - 1): sk_stage<SCALE, PACK, JV, 1>(S, acc, a_off0, a_off1, b_off, c_off0, c_off1); break; \
                case 2 * (JV / 2 - 1) + 1: sk_stage<SCALE, PACK, JV, 2>(S, acc, a_off0, a_off1, b_off, c_off0, c_off1); break;
                switch (variant) {       // warp-uniform
                    GSB_SK_CASE(2) GSB_SK_CASE(4) GSB_SK_CASE(6) GSB_SK_CASE(8)
                    GSB_SK_CASE(10) GSB_SK_CASE(12) GSB_SK_CASE(14) GSB_SK_CASE(16)
                default: break;          // no row group of this warp inside the mesh
                }
#undef GSB_SK_CASE
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[slot]);
            if (++slot == SK_STAGES) { slot = 0; ++round; }
        }

        // ---- segment bookkeeping: an owned tile starts at stage 0 in this CTA ----
        if (s_begin > 0) {
            // leading, non-owned segment: hand the partial accumulators to the owner
            double *dst = prm.slots + (int64_t)cta * (SK_TM * SK_TN) + tid;
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    __stcg(dst + ((i * 16 + j) * 2 + 0) * SK_THREADS, acc[i][j][0]);
                    __stcg(dst + ((i * 16 + j) * 2 + 1) * SK_THREADS, acc[i][j][1]);
                }
            __threadfence();
            __syncthreads();
            if (tid == 0) sk_st_release(prm.flags + cta, 1u);
            continue;
        }
        if (s_end < n_stages) {
            // owned tile that continues in the following CTA(s): add their partials in CTA order
            for (int m = cta + 1; m < (int)gridDim.x; ++m) {
                const SkBound bm = prm.bnd[m], bn = prm.bnd[m + 1];
                if (bm.tile > tile) break;                                     // CTA m starts after this tile
                const bool empty_share = (bm.tile == bn.tile && bm.stage == bn.stage);
                if (!empty_share) {
                    if (tid == 0) {
                        while (sk_ld_acquire(prm.flags + m) == 0u) __nanosleep(64);
                    }
                    __syncthreads();
                    const double *src = prm.slots + (int64_t)m * (SK_TM * SK_TN) + tid;
#pragma unroll
                    for (int i = 0; i < 2; ++i)
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            acc[i][j][0] += __ldcg(src + ((i * 16 + j) * 2 + 0) * SK_THREADS);
                            acc[i][j][1] += __ldcg(src + ((i * 16 + j) * 2 + 1) * SK_THREADS);
                        }
                }
                if (bn.tile > tile) break;                                     // CTA m finished the tile
            }
        }

        // ---- epilogue: thread holds C[g][2t], C[g][2t+1] of every 8x8 tile ----
        const int comp = (int)(z % prm.ncomp);
        double *out = prm.out + z * prm.out_fstride;
        const bool vec2 = (prm.lc & 1) == 0 && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
        const int64_t col0 = (int64_t)ct * SK_TN;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            // virtual row of this thread -> (slow index, row of the tile axis)
            const int64_t v = vt * SK_TM + 8 * (warp + 8 * i) + g;
            if (v >= prm.vrows) continue;
            const int64_t slow = v / prm.lyp;
            const int64_t iy = v - slow * prm.lyp;
            if (iy >= prm.ly) continue;
            const int64_t row = ((slow / prm.n_in) * prm.ly + iy) * prm.n_in + slow % prm.n_in;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int64_t col = col0 + j * 8 + 2 * t;
                const int64_t idx = row * prm.lc + col;
                double *dst = out + idx;
                if (prm.epi.on) {
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

inline int sk_launch(const SkParams &prm, int grid, bool scale, bool partial, bool pack, cudaStream_t st)
{
    static std::atomic<uint64_t> attr_set{0};
    if (first_launch_on_device(attr_set)) {
#define GSB_SK_ATTR(K)                                                                                               \
        GSB_CUDA(cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SK_SMEM_BYTES));
        GSB_SK_ATTR((sk_contract_kernel<false, false, false>)) GSB_SK_ATTR((sk_contract_kernel<false, true, false>))
        GSB_SK_ATTR((sk_contract_kernel<true, false, false>)) GSB_SK_ATTR((sk_contract_kernel<true, true, false>))
        GSB_SK_ATTR((sk_contract_kernel<true, true, true>))
#undef GSB_SK_ATTR
    }
    if (pack) {
        if (!scale) return fail(GSB_ERR_ARGUMENT, "internal: row packing without slow axes");
        sk_contract_kernel<true, true, true><<<grid, SK_THREADS, SK_SMEM_BYTES, st>>>(prm);
    } else if (scale) {
        if (partial) sk_contract_kernel<true, true, false><<<grid, SK_THREADS, SK_SMEM_BYTES, st>>>(prm);
        else sk_contract_kernel<true, false, false><<<grid, SK_THREADS, SK_SMEM_BYTES, st>>>(prm);
    } else {
        if (partial) sk_contract_kernel<false, true, false><<<grid, SK_THREADS, SK_SMEM_BYTES, st>>>(prm);
        else sk_contract_kernel<false, false, false><<<grid, SK_THREADS, SK_SMEM_BYTES, st>>>(prm);
    }
    g_launches.fetch_add(1);
    GSB_CUDA(cudaGetLastError());
    return GSB_OK;
}

}  // namespace gsb
