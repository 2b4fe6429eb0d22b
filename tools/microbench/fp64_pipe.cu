// FP64 pipe microbenchmarks at the occupancy of the separable kernel (few warps per SMSP).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_pipe fp64_pipe.cu
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

// A: register-tiled DFMA, 8x8 accumulators, fragments fixed in registers
template <int T> __global__ void __launch_bounds__(T, 1) dfma_tile(double* sink, int iters, const double* in)
{
    double a[8], b[8], acc[8][8];
    for (int i = 0; i < 8; ++i) { a[i] = in[threadIdx.x + i]; b[i] = in[threadIdx.x + 8 + i]; }
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 8; ++j) acc[i][j] = 0.0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
            // perturb fragments cheaply so the compiler cannot hoist (1 DADD per 64 DFMA)
            a[u] += 1e-30;
        }
    }
    double s = 0; for (int i = 0; i < 8; ++i) for (int j = 0; j < 8; ++j) s += acc[i][j];
    if (s == 1.2345) sink[0] = s;
}

// B: DMMA 4x8 tiles
template <int T> __global__ void __launch_bounds__(T, 1) dmma_tile(double* sink, int iters, const double* in)
{
    double a[4], b[8], acc[4][8][2];
    for (int i = 0; i < 4; ++i) a[i] = in[threadIdx.x + i];
    for (int i = 0; i < 8; ++i) b[i] = in[threadIdx.x + 8 + i];
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 8; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                                 : "+d"(acc[i][j][0]), "+d"(acc[i][j][1]) : "d"(a[i]), "d"(b[j]));
            a[u] += 1e-30;
        }
    }
    double s = 0; for (int i = 0; i < 4; ++i) for (int j = 0; j < 8; ++j) s += acc[i][j][0] + acc[i][j][1];
    if (s == 1.2345) sink[0] = s;
}

// C: simple DFMA chains with constant operands (the "peak" loop), ILP chains
template <int CH>
__global__ void __launch_bounds__(512, 1) dfma_chain(double* sink, int iters, const double* in)
{
    double acc[CH];
    const double m = in[0], c = in[1];
    for (int i = 0; i < CH; ++i) acc[i] = in[threadIdx.x + i];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 256 / CH; ++u)
#pragma unroll
            for (int i = 0; i < CH; ++i) acc[i] = fma(acc[i], m, c);
    }
    double s = 0; for (int i = 0; i < CH; ++i) s += acc[i];
    if (s == 1.2345) sink[0] = s;
}

int main()
{
    double *sink, *in;
    CK(cudaMalloc(&sink, 8));
    CK(cudaMalloc(&in, 8 * 4096));
    { double h[4096]; for (int i = 0; i < 4096; ++i) h[i] = 1.0 + 1e-3 * i; CK(cudaMemcpy(in, h, sizeof h, cudaMemcpyHostToDevice)); }
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4000;
    for (int warps_per_smsp = 1; warps_per_smsp <= 3; ++warps_per_smsp) {
        const int threads = warps_per_smsp * 4 * 32;
        for (int kind = 0; kind < 4; ++kind) {
            double fma_per_thread_iter = 0; const char* name = "";
            float best = 1e30f;
            for (int rep = 0; rep < 4; ++rep) {
                cudaEventRecord(e0);
                switch (kind) {
                case 0: if (threads == 128) dfma_tile<128><<<sms, threads>>>(sink, iters, in); else if (threads == 256) dfma_tile<256><<<sms, threads>>>(sink, iters, in); else dfma_tile<384><<<sms, threads>>>(sink, iters, in); fma_per_thread_iter = 256; name = "DFMA 8x8 tile"; break;
                case 1: if (threads == 128) dmma_tile<128><<<sms, threads>>>(sink, iters, in); else if (threads == 256) dmma_tile<256><<<sms, threads>>>(sink, iters, in); else dmma_tile<384><<<sms, threads>>>(sink, iters, in); fma_per_thread_iter = 4 * 32 * 8; name = "DMMA 4x8 tiles"; break;
                case 2: dfma_chain<8><<<sms, threads>>>(sink, iters, in); fma_per_thread_iter = 256; name = "DFMA chain x8"; break;
                case 3: dfma_chain<32><<<sms, threads>>>(sink, iters, in); fma_per_thread_iter = 256; name = "DFMA chain x32"; break;
                }
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms;
            }
            CK(cudaGetLastError());
            const double fmas = (double)sms * threads * iters * fma_per_thread_iter;
            printf("warps/SMSP=%d %-16s %.3f TFMA/s (%.1f%% of 148*64*1.965e9)\n", warps_per_smsp, name,
                   fmas / (best * 1e-3) / 1e12, 100.0 * fmas / (best * 1e-3) / (148.0 * 64 * 1.965e9));
        }
    }
    return 0;
}
