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
// ~D+17 of the direct kernel: all sin/cos work moves into per-axis tables of size (len_t x N)
// built once per call with full-accuracy sincos.
//
// Kernel structure (one CTA per 128x128 output tile, 1 CTA/SM, warp specialised):
//   * warps 8..11 = TMA warpgroup: one elected thread issues, per pipeline stage of KC modes, the
//     cp.async.bulk copies (TMA unit) of the Cz/Sz table slices into shared memory, signalled on
//     the stage's "full" mbarrier; the warpgroup gives its registers to the consumers
//     (setmaxnreg.dec / .inc).
//   * the A operand is GENERATED on the fly by the consumer threads, one stage ahead: one complex
//     product per (row, mode) from the L2-resident row-axis tables, written straight to shared
//     memory.  The 4.2 GB A matrix of config 2 never exists in HBM.
//   * warps 0..7 = consumers: each warp owns a 32x64 block of the tile as 4x8 DMMA.8x8x4
//     accumulator tiles (64 fp64 accumulators per thread); per 2 modes it fetches 4 + 8 fragment
//     doubles with conflict-free LDS.64 and issues 32 DMMA; modes are accumulated in ascending
//     order (deterministic, no atomics).  They release the stage on its "empty" mbarrier.
//   * epilogue: registers -> global, 16-byte stores.
// The incompressible variant (generator.py:479-495) multiplies A by the projector p_t(k_j) and
// runs one CTA column per vector component (2 d DFMA per pair).
#pragma once

#include "gsb_common.cuh"

namespace gsb {

constexpr int SEP_TM = 128;      // rows per CTA tile
constexpr int SEP_TN = 128;      // columns per CTA tile
#ifndef GSB_SEP_KC
#define GSB_SEP_KC 8
#endif
#ifndef GSB_SEP_STAGES
#define GSB_SEP_STAGES 5
#endif
#ifndef GSB_SEP_LOOKAHEAD
#define GSB_SEP_LOOKAHEAD 2
#endif
constexpr int SEP_KC = GSB_SEP_KC;        // modes per pipeline stage (multiple of 4)
constexpr int SEP_STAGES = GSB_SEP_STAGES;
constexpr int SEP_LOOKAHEAD = GSB_SEP_LOOKAHEAD;   // A operand is generated this many stages ahead (needs STAGES >= 2*LOOKAHEAD)
static_assert(SEP_KC % 4 == 0 && SEP_STAGES >= 2 * SEP_LOOKAHEAD + 1, "pipeline shape");
constexpr int SEP_MAX_ROW_AXES = GSB_MAX_DIM - 1;


struct SepParams {
    // row-axis tables: E_t[j * len_t + i] = exp(i k'_{t,j} a_t[i]) as (cos, sin); t = 0 is
    // pre-multiplied by (z1_j - i z2_j).  Modes j >= n_modes (padding) are zero.
    const double2 *erow[SEP_MAX_ROW_AXES];
    int64_t row_len[SEP_MAX_ROW_AXES];     // extent of each row axis in THIS launch (slab)
    int64_t row_stride[SEP_MAX_ROW_AXES];  // full table width (entries per mode) of each row axis
    int n_row_axes;
    int64_t n_rows;        // prod(row_len)
    // last-axis tables, planes of n_modes_pad x lc_pad doubles
    const double *bc;
    const double *bs;
    int64_t lc;            // length of the last axis
    int64_t lc_pad;        // padded to a multiple of SEP_TN
    int n_modes_pad;       // multiple of SEP_KC
    // incompressible projector p_t[j], (ncomp, n_modes_pad); nullptr for the scalar sum
    const double *proj;
    int ncomp;             // 1 (scalar) or dim (vector)
    // per-batch strides (in elements) of the tables above; batch index = blockIdx.z / ncomp
    int64_t erow_bstride[SEP_MAX_ROW_AXES];
    int64_t b_bstride;
    int64_t proj_bstride;
    double *out;           // field f = batch*ncomp + comp starts at out + f*out_fstride; (n_rows, lc) inside
    int64_t out_fstride;
};

// ---------------------------------------------------------------------------------------------
// table builder: one thread per (axis entry, mode).  Full-accuracy sincos (libdevice), cost
// O((sum_t len_t) N) -- negligible against the O(n N) contraction.
// ---------------------------------------------------------------------------------------------
struct TableParams {
    const double *cov;     // (n_batch, dim, n_modes)
    const double *z1;      // (n_batch, n_modes)
    const double *z2;
    const double *axes;    // concatenated axis coordinates
    int64_t axis_off[GSB_MAX_DIM];
    int64_t axis_len[GSB_MAX_DIM];
    double matrix[GSB_MAX_DIM * GSB_MAX_DIM];  // row-major (dim x dim) isometrisation matrix
    int dim;
    int64_t n_modes;
    int n_modes_pad;
    int vec;               // build the projector table
    double2 *erow[SEP_MAX_ROW_AXES];
    int64_t erow_bstride[SEP_MAX_ROW_AXES];
    double *bc;
    double *bs;
    int64_t lc_pad;
    int64_t b_bstride;
    double *proj;
    int64_t proj_bstride;
};

__global__ void build_tables_kernel(const TableParams tp)
{
    const int t = blockIdx.y;                 // axis
    const int64_t b = blockIdx.z;             // batch entry
    const int64_t len = tp.axis_len[t];
    const bool last = (t == tp.dim - 1);
    const int64_t width = last ? tp.lc_pad : len;
    const int64_t total = width * tp.n_modes_pad;
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
            sincos(kp * tp.axes[tp.axis_off[t] + i], &s, &c);
            if (t == 0 && !last) {  // fold the complex weight (z1 - i z2) into the first row axis
                const double a = tp.z1[b * tp.n_modes + j], bb = tp.z2[b * tp.n_modes + j];
                const double re = a * c + bb * s;
                const double im = a * s - bb * c;
                c = re;
                s = im;
            }
        }
        if (last) {
            tp.bc[b * tp.b_bstride + j * tp.lc_pad + i] = c;
            tp.bs[b * tp.b_bstride + j * tp.lc_pad + i] = s;
        } else {
            tp.erow[t][b * tp.erow_bstride[t] + j * len + i] = make_double2(c, s);
        }
    }
    // projector table (vector field): p_c[j] = delta_c0 - k_c k_0 / |k|^2 on the ORIGINAL k
    if (tp.vec && t == 0) {
        for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < tp.n_modes_pad;
             j += (int64_t)gridDim.x * blockDim.x) {
            double k2 = 0.0;
            if (j < tp.n_modes)
                for (int s2 = 0; s2 < tp.dim; ++s2) {
                    const double k = cov[(int64_t)s2 * tp.n_modes + j];
                    k2 += k * k;
                }
            for (int c2 = 0; c2 < tp.dim; ++c2) {
                double p = 0.0;
                if (j < tp.n_modes) {
                    const double e = (c2 == 0) ? 1.0 : 0.0;
                    p = e - cov[(int64_t)c2 * tp.n_modes + j] * cov[j] / k2;
                }
                tp.proj[b * tp.proj_bstride + (int64_t)c2 * tp.n_modes_pad + j] = p;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// the contraction kernel
//
// The consumers run DMMA.8x8x4 (mma.sync m8n8k4 f64) on the tensor-core FP64 path.  B200 executes
// it at the same FMA rate as DFMA (measured 18.5 vs 18.4 TFMA/s), but one instruction carries 256
// FMAs with 8 register reads, so neither the register file (3 x 64-bit operands per DFMA exceed
// the 2 reads/clk it sustains once the operand-reuse cache is lost between warps: a register-tiled
// DFMA version of this kernel measured 73 % of the FP64 peak, profiles/) nor shared-memory
// fragment traffic limits the pipe.
// Shared-memory layout per stage:
//   A[row][SEP_AST]           k = 2*kc + part (part 0: p*Re A, part 1: -p*Im A); rows padded to
//                             2*KC+4 doubles so the 8x4 fragment loads (LDS.64) hit 32 distinct
//                             banks per half warp;
//   B[2*kc + part][SEP_BST]   part 0: cos, part 1: sin; rows padded to 132 doubles likewise.
// Two warp configurations (template CFG), no warp specialisation (every warp contracts, every
// thread generates part of A, thread 0 issues the TMA copies):
//   CFG 0: 16 warps (4 per SM sub-partition), warp tile 32x32 = 4x4 DMMA tiles, 32 accumulators
//          per thread, 512 threads x 128 registers;
//   CFG 1:  8 warps (2 per sub-partition), warp tile 32x64 = 4x8 DMMA tiles, 64 accumulators per
//          thread, 256 threads x 255 registers.
// ---------------------------------------------------------------------------------------------
constexpr int SEP_AST = 2 * SEP_KC + 4;   // A row stride (doubles)
constexpr int SEP_BST = SEP_TN + 4;       // B row stride (doubles)
constexpr int SEP_STAGE_DOUBLES = SEP_TM * SEP_AST + 2 * SEP_KC * SEP_BST;
constexpr size_t SEP_SMEM_BYTES =
    (size_t)SEP_STAGES * SEP_STAGE_DOUBLES * sizeof(double) + SEP_STAGES * sizeof(uint64_t) + 128;

template <int CFG>
struct SepCfg;
template <>
struct SepCfg<0> {
    static constexpr int WARPS = 16, WCOLS = 4, RT = 4, CT = 4;   // 512 threads x 128 registers
};
template <>
struct SepCfg<1> {
    static constexpr int WARPS = 8, WCOLS = 2, RT = 4, CT = 8;    // 256 threads x 255 registers
};

__device__ __forceinline__ void dmma_884(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

template <int NRA, int CFG>  // NRA = number of row axes (dim - 1), 1..SEP_MAX_ROW_AXES
__device__ __forceinline__ void separable_body(const SepParams &prm)
{
    using C = SepCfg<CFG>;
    constexpr int NTHR = C::WARPS * 32;
    constexpr int RT = C::RT, CT = C::CT;
    constexpr int GEN = SEP_TM * SEP_KC / NTHR;   // A entries generated per thread per stage
    static_assert(C::WARPS / C::WCOLS * RT * 8 == SEP_TM && C::WCOLS * CT * 8 == SEP_TN, "tile");
    static_assert(GEN >= 1 && SEP_TM * SEP_KC % NTHR == 0, "A generation split");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *stage_base = reinterpret_cast<double *>(smem_raw);
    uint64_t *full = reinterpret_cast<uint64_t *>(stage_base + SEP_STAGES * SEP_STAGE_DOUBLES);

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    const int64_t col0 = (int64_t)blockIdx.x * SEP_TN;
    const int64_t row0 = (int64_t)blockIdx.y * SEP_TM;
    const int comp = blockIdx.z % prm.ncomp;
    const int64_t batch = blockIdx.z / prm.ncomp;
    const int n_stages_total = prm.n_modes_pad / SEP_KC;

    if (tid == 0) {
        for (int s = 0; s < SEP_STAGES; ++s) mbar_init(&full[s], NTHR + 1);  // A writers + TMA issuer
        fence_barrier_init();
    }
    __syncthreads();

    const int wr = warp / C::WCOLS;      // row band of RT*8 rows
    const int wc = warp % C::WCOLS;      // column band of CT*8 columns
    double *out = prm.out + (batch * prm.ncomp + comp) * prm.out_fstride;
    const bool vec2 = (prm.lc & 1) == 0 && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);

    // ---- B operand: thread 0 issues the cp.async.bulk copies (TMA unit) of stage s -------------
    // 2*KC table-row slices of 1 KB each, completion counted on the stage's mbarrier.
    const double *bc = prm.bc + batch * prm.b_bstride + col0;
    const double *bs = prm.bs + batch * prm.b_bstride + col0;
    auto tma_issue = [&](int s) {
        const int slot = s % SEP_STAGES;
        double *B = stage_base + slot * SEP_STAGE_DOUBLES + SEP_TM * SEP_AST;
        const int64_t j0 = (int64_t)s * SEP_KC;
        mbar_arrive_expect_tx(&full[slot], 2 * SEP_KC * SEP_TN * sizeof(double));
#pragma unroll
        for (int kc = 0; kc < SEP_KC; ++kc) {
            bulk_g2s(B + (2 * kc) * SEP_BST, bc + (j0 + kc) * prm.lc_pad, SEP_TN * sizeof(double), &full[slot]);
            bulk_g2s(B + (2 * kc + 1) * SEP_BST, bs + (j0 + kc) * prm.lc_pad, SEP_TN * sizeof(double), &full[slot]);
        }
    };

    // ---- A operand: generated cooperatively, thread -> (tile row, GEN of the KC modes) ---------
    // The loads for stage s+LOOKAHEAD are issued before the contraction of stage s and consumed
    // after it, so their L2 latency is hidden; the complex products run on the issuing warp's own
    // FP64 slots (in order with its DMMAs -- a dedicated producer warp starves behind them).
    // LOOKAHEAD = 2 keeps the warps out of lock step: full[s+1] only needs every warp to have
    // finished stage s-1.
    // Slot reuse needs no "empty" barrier when STAGES >= 2*LOOKAHEAD + 1: a thread that has
    // passed full[s] knows that ALL threads arrived on it, which each does only after finishing
    // the contraction of stage s-LOOKAHEAD; the slot of stage s+LOOKAHEAD was last read by stage
    // s+LOOKAHEAD-STAGES <= s-LOOKAHEAD-1.  Both the A writes and the TMA issue for stage
    // s+LOOKAHEAD happen after that wait.
    const int grow = tid & (SEP_TM - 1);
    const int gm0 = (tid / SEP_TM) * GEN;
    const double2 *ep[NRA];
    {
        int64_t r = row0 + grow;
        if (r >= prm.n_rows) r = prm.n_rows - 1;  // clamp; result is never stored
#pragma unroll
        for (int t = NRA - 1; t >= 0; --t) {
            const int64_t it = r % prm.row_len[t];
            r /= prm.row_len[t];
            ep[t] = prm.erow[t] + batch * prm.erow_bstride[t] + it;
        }
    }
    const double *proj =
        prm.proj ? prm.proj + batch * prm.proj_bstride + (int64_t)comp * prm.n_modes_pad : nullptr;
    double2 ge[NRA][GEN];
    auto gen_load = [&](int s) {
        const int64_t j = (int64_t)s * SEP_KC + gm0;
#pragma unroll
        for (int t = 0; t < NRA; ++t)
#pragma unroll
            for (int u = 0; u < GEN; ++u) ge[t][u] = __ldg(ep[t] + (j + u) * prm.row_stride[t]);
    };
    auto gen_store = [&](int s) {
        const int slot = s % SEP_STAGES;
        double *A = stage_base + slot * SEP_STAGE_DOUBLES;
#pragma unroll
        for (int u = 0; u < GEN; ++u) {
            double2 e = ge[0][u];
#pragma unroll
            for (int t = 1; t < NRA; ++t) {
                const double re = e.x * ge[t][u].x - e.y * ge[t][u].y;
                const double im = e.x * ge[t][u].y + e.y * ge[t][u].x;
                e.x = re;
                e.y = im;
            }
            if (proj) {
                const double pj = proj[(int64_t)s * SEP_KC + gm0 + u];
                e.x *= pj;
                e.y *= pj;
            }
            *reinterpret_cast<double2 *>(A + grow * SEP_AST + 2 * (gm0 + u)) = make_double2(e.x, -e.y);
        }
        mbar_arrive(&full[slot]);  // release: this thread's part of the A tile is written
    };

#pragma unroll
    for (int p = 0; p < SEP_LOOKAHEAD; ++p) {
        if (p < n_stages_total) {
            if (tid == 0) tma_issue(p);
            gen_load(p);
            gen_store(p);
        }
    }

    // warp tile = RT x CT DMMA tiles of 8x8; fragment owner: g = lane/4, t = lane%4
    const int g = lane >> 2;
    const int t = lane & 3;
    double acc[RT][CT][2];
#pragma unroll
    for (int i = 0; i < RT; ++i)
#pragma unroll
        for (int j = 0; j < CT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    const int a_off = (wr * RT * 8 + g) * SEP_AST + t;   // + i*8*SEP_AST + 4*k4
    const int b_off = t * SEP_BST + wc * CT * 8 + g;     // + 4*k4*SEP_BST + j*8

    for (int s = 0; s < n_stages_total; ++s) {
        const int slot = s % SEP_STAGES;
        if (s + SEP_LOOKAHEAD < n_stages_total) gen_load(s + SEP_LOOKAHEAD);
        mbar_wait(&full[slot], (s / SEP_STAGES) & 1);
        if (tid == 0 && s + SEP_LOOKAHEAD < n_stages_total) tma_issue(s + SEP_LOOKAHEAD);
        const double *A = stage_base + slot * SEP_STAGE_DOUBLES;
        const double *B = A + SEP_TM * SEP_AST;
#pragma unroll
        for (int k4 = 0; k4 < SEP_KC / 2; ++k4) {   // 4 contraction indices = 2 modes
            double af[RT], bf[CT];
#pragma unroll
            for (int i = 0; i < RT; ++i) af[i] = A[a_off + i * 8 * SEP_AST + 4 * k4];
#pragma unroll
            for (int j = 0; j < CT; ++j) bf[j] = B[b_off + 4 * k4 * SEP_BST + j * 8];
#pragma unroll
            for (int j = 0; j < CT; ++j)
#pragma unroll
                for (int i = 0; i < RT; ++i) dmma_884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
        if (s + SEP_LOOKAHEAD < n_stages_total) gen_store(s + SEP_LOOKAHEAD);
    }

    // epilogue: thread holds C[g][2t], C[g][2t+1] of every 8x8 tile
#pragma unroll
    for (int i = 0; i < RT; ++i) {
        const int64_t row = row0 + wr * RT * 8 + i * 8 + g;
        if (row >= prm.n_rows) continue;
#pragma unroll
        for (int j = 0; j < CT; ++j) {
            const int64_t col = col0 + wc * CT * 8 + j * 8 + 2 * t;
            double *dst = out + row * prm.lc + col;
            if (vec2 && col + 1 < prm.lc) {
                *reinterpret_cast<double2 *>(dst) = make_double2(acc[i][j][0], acc[i][j][1]);
            } else {
                if (col < prm.lc) dst[0] = acc[i][j][0];
                if (col + 1 < prm.lc) dst[1] = acc[i][j][1];
            }
        }
    }
}

// CFG 0: 16 warps x 128 registers
template <int NRA>
__global__ void __launch_bounds__(512, 1) separable_kernel_w16(const SepParams prm)
{
    separable_body<NRA, 0>(prm);
}

// CFG 1: 8 warps x 255 registers
template <int NRA>
__global__ void __launch_bounds__(256, 1) separable_kernel_w8(const SepParams prm)
{
    separable_body<NRA, 1>(prm);
}

template <int NRA, int CFG>
inline int launch_separable_variant(const SepParams &prm, dim3 grid, cudaStream_t st)
{
    using C = SepCfg<CFG>;
    constexpr int threads = C::WARPS * 32;
    if (CFG == 0) {
        GSB_CUDA(cudaFuncSetAttribute(separable_kernel_w16<NRA>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)SEP_SMEM_BYTES));
        separable_kernel_w16<NRA><<<grid, threads, SEP_SMEM_BYTES, st>>>(prm);
    } else {
        GSB_CUDA(cudaFuncSetAttribute(separable_kernel_w8<NRA>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)SEP_SMEM_BYTES));
        separable_kernel_w8<NRA><<<grid, threads, SEP_SMEM_BYTES, st>>>(prm);
    }
    return GSB_OK;
}

// variant: warp configuration CFG (0 = 16 consumer warps, 1 = 8 consumer warps + setmaxnreg)
inline int launch_separable(const SepParams &prm, int64_t n_batch, int variant, cudaStream_t st)
{
    dim3 grid((unsigned)(prm.lc_pad / SEP_TN), (unsigned)((prm.n_rows + SEP_TM - 1) / SEP_TM),
              (unsigned)(n_batch * prm.ncomp));
    if (grid.y > 65535u || grid.z > 65535u)
        return fail(GSB_ERR_ARGUMENT, "structured mesh too large for one launch (rows/128 or batch*ncomp > 65535)");
#define GSB_SEP_CASE(N)                                                                        \
    case N:                                                                                    \
        if (variant == 1) GSB_TRY((launch_separable_variant<N, 1>(prm, grid, st)));            \
        else GSB_TRY((launch_separable_variant<N, 0>(prm, grid, st)));                         \
        break;
    switch (prm.n_row_axes) {
        GSB_SEP_CASE(1)
        GSB_SEP_CASE(2)
        GSB_SEP_CASE(3)
        GSB_SEP_CASE(4)
        GSB_SEP_CASE(5)
        GSB_SEP_CASE(6)
        GSB_SEP_CASE(7)
    default:
        return fail(GSB_ERR_ARGUMENT, "structured path needs 2 <= dim <= 8");
    }
#undef GSB_SEP_CASE
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
