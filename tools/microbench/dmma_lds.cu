// DMMA consumer loop microbenchmark: the separable kernel's warp tile (32x64 = 4x8 DMMA tiles)
// fed from shared memory with the kernel's fragment layout, without barriers / A generation.
// Variants: WT=0 warp tile 32x64 (8 warps), WT=1 warp tile 32x32 (16 warps).
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

constexpr int KC = 8, AST = 2 * KC + 4, BST = 132, TM = 128;

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NT, int RT, int CT>   // NT threads, warp tile = RT x CT dmma tiles
__global__ void __launch_bounds__(NT, 1) loop(double* sink, int stages)
{
    extern __shared__ double sm[];
    double* A = sm;                 // [128][AST]
    double* B = sm + TM * AST;      // [2*KC][BST]
    for (int i = threadIdx.x; i < TM * AST + 2 * KC * BST; i += NT) sm[i] = 1.0 + 1e-6 * i;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    constexpr int WCOLS = 128 / (CT * 8);           // warps along columns
    const int wr = warp / WCOLS, wc = warp % WCOLS;
    const int a_off = (wr * RT * 8 + g) * AST + t;
    const int b_off = t * BST + wc * CT * 8 + g;
    double acc[RT][CT][2];
#pragma unroll
    for (int i = 0; i < RT; ++i)
#pragma unroll
        for (int j = 0; j < CT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    for (int s = 0; s < stages; ++s) {
#pragma unroll
        for (int k4 = 0; k4 < KC / 2; ++k4) {
            double af[RT], bf[CT];
#pragma unroll
            for (int i = 0; i < RT; ++i) af[i] = A[a_off + i * 8 * AST + 4 * k4];
#pragma unroll
            for (int j = 0; j < CT; ++j) bf[j] = B[b_off + 4 * k4 * BST + j * 8];
#pragma unroll
            for (int i = 0; i < RT; ++i)
#pragma unroll
                for (int j = 0; j < CT; ++j) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
    }
    double r = 0;
#pragma unroll
    for (int i = 0; i < RT; ++i)
#pragma unroll
        for (int j = 0; j < CT; ++j) r += acc[i][j][0] + acc[i][j][1];
    if (r == 1.2345) sink[0] = r;
}

template <int NT, int RT, int CT>
int run(const char* name, double* sink, int sms)
{
    const int stages = 2000;
    const size_t smem = (TM * AST + 2 * KC * BST) * sizeof(double);
    CK(cudaFuncSetAttribute(loop<NT, RT, CT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        loop<NT, RT, CT><<<sms, NT, smem>>>(sink, stages);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms;
    }
    CK(cudaGetLastError());
    const double fmas = (double)sms * (NT / 32) * stages * (KC / 2) * RT * CT * 256.0;
    printf("%-34s %.3f TFMA/s (%.1f%% of 148*64*1.965e9)\n", name, fmas / (best * 1e-3) / 1e12,
           100.0 * fmas / (best * 1e-3) / (148.0 * 64 * 1.965e9));
    return 0;
}

int main()
{
    double* sink; CK(cudaMalloc(&sink, 8));
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    run<256, 4, 8>("8 warps, warp tile 32x64 (4x8)", sink, p.multiProcessorCount);
    run<512, 4, 4>("16 warps, warp tile 32x32 (4x4)", sink, p.multiProcessorCount);
    run<512, 2, 8>("16 warps, warp tile 16x64 (2x8)", sink, p.multiProcessorCount);
    run<384, 4, 8>("12 warps(!), warp tile 32x64", sink, p.multiProcessorCount);
    run<128, 4, 8>("4 warps, warp tile 32x64", sink, p.multiProcessorCount);
    return 0;
}
