// gsb_separable.cuh -- shared pieces of the tiled fp64 contractions for sm_100a.
//
// The separable summation on structured meshes lives in gsb_sepk.cuh (second generation: no pre-generated A
// operand, stream-K work split).  The FIRST generation -- per-axis tables, an A-operand generator (agen_kernel)
// writing the pre-tiled operand to HBM from a helper stream, and a contraction walking whole 128 x 128 tiles in
// row chunks -- was removed in round 2: at 512^3 both run at 93-94 % of the FP64 peak, but the first generation
// moved ~12 GB through HBM per field, needed 3 GiB of scratch, and fell to 27-60 % on meshes that do not fill
// whole waves of tiles (profiles/r01_odd_mesh_bench.log vs profiles/r02_odd_mesh_bench.log).
// What is left here is what the kriging contraction (gsb_krige.cuh) shares with it: the 128 x 128 x (8 modes)
// pipeline-stage layout, the DMMA.8x8x4 wrapper, and the device-side mesh expansion for the direct kernel.
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
constexpr int SEP_KC = 8;        // depth pairs per pipeline stage (16 contraction indices)
constexpr int SEP_WARPS = 8;
constexpr int SEP_THREADS = SEP_WARPS * 32;

// Stage layout in shared memory == tile layout in global memory (doubles):
//   A tile [SEP_TM rows][SEP_AST]   the row stride 2*KC+4 makes the 8x4 fragment loads (LDS.64) hit 32 distinct
//                                   banks per half warp
//   B tile [2*KC][SEP_BST]          stride 132
constexpr int SEP_AST = 2 * SEP_KC + 4;
constexpr int SEP_BST = SEP_TN + 4;
constexpr int SEP_A_TILE = SEP_TM * SEP_AST;          // doubles
constexpr int SEP_B_TILE = 2 * SEP_KC * SEP_BST;      // doubles
static_assert((SEP_A_TILE * 8) % 16 == 0 && (SEP_B_TILE * 8) % 16 == 0, "bulk copy granularity");

__device__ __forceinline__ void dmma_884(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// device-side mesh expansion for meshes that go through the direct kernel:
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
