// Pure TMA-fed DMMA contraction: one producer thread issues 2 bulk copies per stage (a 20 KB A
// tile and a 16.5 KB B tile, both pre-tiled in global memory), 8 consumer warps run DMMA.8x8x4
// on 32x64 warp tiles, full/empty mbarrier ring.  Upper bound for the two-kernel design.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

constexpr int KC = 8, AST = 2 * KC + 4, BST = 132, TM = 128;
constexpr int A_D = TM * AST, B_D = 2 * KC * BST, STAGE_D = A_D + B_D;

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n"
                 ::"r"(s32(bar)), "r"(parity) : "memory");
}

template <int STAGES, int CW, bool SCALE, int WT = 0>   // CW consumer warps: 8 (32x64 tiles) or 16 (32x32 tiles)
__global__ void __launch_bounds__((CW + 1) * 32, 1) gemm_like(double* sink, int stages, const double* at, const double* bt, int n_at)
{
    extern __shared__ __align__(128) double sm[];
    uint64_t* full = reinterpret_cast<uint64_t*>(sm + STAGES * STAGE_D);
    uint64_t* empty = full + STAGES;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&full[s])), "r"(1));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&empty[s])), "r"(CW));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (warp == CW) {
        if (lane != 0) return;
        const double* a = at + (size_t)(blockIdx.x % n_at) * A_D * 125;
        for (int s = 0; s < stages; ++s) {
            const int slot = s % STAGES, round = s / STAGES;
            if (round > 0) mbar_wait(&empty[slot], (round - 1) & 1);
            double* A = sm + slot * STAGE_D;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[slot])), "r"((A_D + B_D) * 8) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(s32(A)), "l"(a + (size_t)(s % 125) * A_D), "r"(A_D * 8), "r"(s32(&full[slot])) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(s32(A + A_D)), "l"(bt + (size_t)(s % 125) * B_D), "r"(B_D * 8), "r"(s32(&full[slot])) : "memory");
        }
        return;
    }
    constexpr int WCOLS = WT ? 1 : ((CW == 8) ? 2 : 4), CT = WT ? 16 : ((CW == 8) ? 8 : 4), RT = WT ? 2 : 4;
    const int g = lane >> 2, t = lane & 3, wr = warp / WCOLS, wc = warp % WCOLS;
    const int a_off = (wr * RT * 8 + g) * AST + t, b_off = t * BST + wc * CT * 8 + g;
    double acc[RT][CT][2];
#pragma unroll
    for (int i = 0; i < RT; ++i)
#pragma unroll
        for (int j = 0; j < CT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    for (int s = 0; s < stages; ++s) {
        const int slot = s % STAGES;
        mbar_wait(&full[slot], (s / STAGES) & 1);
        const double* A = sm + slot * STAGE_D;
        const double* B = A + A_D;
#pragma unroll
        for (int k4 = 0; k4 < KC / 2; ++k4) {
            double af[RT], bf[CT];
            if (!SCALE) {
#pragma unroll
                for (int i = 0; i < RT; ++i) af[i] = A[a_off + i * 8 * AST + 4 * k4];
            }
            if (SCALE) {
                // frag' = alpha*own + beta*partner, (alpha, beta) per (lane parity, mode): the slow-axis
                // phase factor applied in the consumer instead of a pre-generated A operand
                const double2 cc = *reinterpret_cast<const double2*>(B + 2 * KC * BST - 32 + 2 * (2 * k4 + (t >> 1)));
                const double alpha = (t & 1) ? -cc.x : cc.x, beta = -cc.y;
#pragma unroll
                for (int i = 0; i < RT; ++i) {
                    // (re, im) pair of this lane's mode: one 16-byte load instead of LDS.64 + shuffle
                    const double2 p = *reinterpret_cast<const double2*>(A + a_off - t + 2 * (t >> 1) + i * 8 * AST + 4 * k4);
                    const double own = (t & 1) ? p.y : p.x, partner = (t & 1) ? p.x : p.y;
                    af[i] = fma(alpha, own, beta * partner);
                }
            }
#pragma unroll
            for (int j = 0; j < CT; ++j) bf[j] = B[b_off + 4 * k4 * BST + j * 8];
#pragma unroll
            for (int j = 0; j < CT; ++j)
#pragma unroll
                for (int i = 0; i < RT; ++i) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&empty[slot])) : "memory");
    }
    double r = 0;
#pragma unroll
    for (int i = 0; i < RT; ++i)
#pragma unroll
        for (int j = 0; j < CT; ++j) r += acc[i][j][0] + acc[i][j][1];
    if (r == 1.2345) sink[0] = r;
}

template <int STAGES, int CW, bool SCALE, int WT = 0>
int run(const char* name, double* sink, int ctas, const double* at, const double* bt, int n_at)
{
    const int stages = 125 * 8;
    const size_t smem = STAGES * STAGE_D * sizeof(double) + 2 * STAGES * 8 + 64;
    CK(cudaFuncSetAttribute(gemm_like<STAGES, CW, SCALE, WT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        gemm_like<STAGES, CW, SCALE, WT><<<ctas, (CW + 1) * 32, smem>>>(sink, stages, at, bt, n_at);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms;
    }
    CK(cudaGetLastError());
    const double fmas = (double)ctas * stages * 8.0 * 128 * 128 * 2;
    printf("%-44s %.3f TFMA/s (%.1f%%)  %.2f ms\n", name, fmas / (best * 1e-3) / 1e12,
           100.0 * fmas / (best * 1e-3) / (148.0 * 64 * 1.965e9), best);
    return 0;
}

int main()
{
    double* sink; CK(cudaMalloc(&sink, 64));
    const int n_at = 2048;   // distinct A row tiles: 2048 * 125 * 20 KB = 5.2 GB, as config 2
    double *at, *bt;
    CK(cudaMalloc(&at, sizeof(double) * (size_t)A_D * 125 * n_at)); CK(cudaMemset(at, 0, sizeof(double) * (size_t)A_D * 125 * n_at));
    CK(cudaMalloc(&bt, sizeof(double) * (size_t)B_D * 125)); CK(cudaMemset(bt, 0, sizeof(double) * (size_t)B_D * 125));
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int n = p.multiProcessorCount;
    run<4, 8, false>("8 consumer warps, 4 stages, 1 wave", sink, n, at, bt, n_at);
    run<5, 8, false>("8 consumer warps, 5 stages, 1 wave", sink, n, at, bt, n_at);
    run<5, 16, false>("16 consumer warps, 5 stages, 1 wave", sink, n, at, bt, n_at);
    run<5, 8, false>("8 consumer warps, 5 stages, 4 waves", sink, 4 * n, at, bt, n_at);
    run<4, 8, true>("8 warps, 4 stages, A rescaled in consumer", sink, n, at, bt, n_at);
    run<5, 8, true>("8 warps, 5 stages, A rescaled in consumer", sink, n, at, bt, n_at);
    run<5, 16, true>("16 warps, 5 stages, A rescaled in consumer", sink, n, at, bt, n_at);
    run<4, 8, false, 1>("8 warps, warp tile 16x128, plain", sink, n, at, bt, n_at);
    run<4, 8, true, 1>("8 warps, warp tile 16x128, A rescaled", sink, n, at, bt, n_at);
    return 0;
}
