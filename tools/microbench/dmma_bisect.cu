// Bisecting what costs the separable kernel its last ~12 %: the pure DMMA+LDS loop reaches 99 % of
// the FP64 peak; add the kernel's other per-stage ingredients one at a time.
//   F1: 16 FP64-pipe ops per thread per stage (the complex products of the A generation)
//   F2: 8 LDG.128 per thread per stage from an L2-resident table
//   F4: 4 STS.128 per thread + mbarrier arrive / wait (CTA-wide) per stage
//   F8: 16 cp.async.bulk copies of 1 KB per stage by thread 0, counted on the same mbarrier
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

constexpr int KC = 8, AST = 2 * KC + 4, BST = 132, TM = 128, STAGES = 5;
constexpr int STAGE_D = TM * AST + 2 * KC * BST;

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

template <int F>
__global__ void __launch_bounds__(256, 1) loop(double* sink, int stages, const double2* tab, const double* btab)
{
    extern __shared__ __align__(128) double sm[];
    uint64_t* full = reinterpret_cast<uint64_t*>(sm + STAGES * STAGE_D);
    for (int i = threadIdx.x; i < STAGES * STAGE_D; i += 256) sm[i] = 1.0 + 1e-6 * i;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&full[s])), "r"(256 + ((F & 8) ? 1 : 0)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int wr = warp >> 1, wc = warp & 1;
    const int a_off = (wr * 32 + g) * AST + t, b_off = t * BST + wc * 64 + g;
    const int grow = tid & 127, gm0 = (tid >> 7) * 4;
    double acc[4][8][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    double2 ge[2][4];
#pragma unroll
    for (int u = 0; u < 4; ++u) { ge[0][u] = make_double2(1.0 + tid, 0.5); ge[1][u] = make_double2(0.25, 2.0 + u); }

    auto tma = [&](int s) {
        const int slot = s % STAGES;
        double* B = sm + slot * STAGE_D + TM * AST;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[slot])), "r"(16 * 1024) : "memory");
#pragma unroll
        for (int k = 0; k < 16; ++k)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(s32(B + k * BST)), "l"(btab + (size_t)((s * 16 + k) % 2000) * 512), "r"(1024), "r"(s32(&full[slot])) : "memory");
    };
    // split phase (F & 16): products are computed at the TOP of a stage from the raw (TMA-loaded)
    // operands and only stored at the END of the stage, a whole stage of DMMAs later
    double2 sp[4];
    auto gen_compute = [&](int s) {
        const int slot = s % STAGES;
        const double* raw = sm + slot * STAGE_D;      // stands in for the raw table tile of stage s
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const double2 e = *reinterpret_cast<const double2*>(raw + grow * AST + 2 * (gm0 + u));
            const double2 c = *reinterpret_cast<const double2*>(raw + TM * AST + 2 * (gm0 + u));   // broadcast
            sp[u].x = e.x * c.x - e.y * c.y;
            sp[u].y = -(e.x * c.y + e.y * c.x);
        }
    };
    auto gen_commit = [&](int s) {
        const int slot = s % STAGES;
        double* A = sm + slot * STAGE_D;
#pragma unroll
        for (int u = 0; u < 4; ++u) *reinterpret_cast<double2*>(A + grow * AST + 2 * (gm0 + u)) = sp[u];
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&full[slot])) : "memory");
    };
    auto gen_store = [&](int s) {
        const int slot = s % STAGES;
        double* A = sm + slot * STAGE_D;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            double2 e = ge[0][u];
            if (F & 1) {
                const double re = e.x * ge[1][u].x - e.y * ge[1][u].y;
                const double im = e.x * ge[1][u].y + e.y * ge[1][u].x;
                e.x = re; e.y = im;
            }
            if (F & 4) *reinterpret_cast<double2*>(A + grow * AST + 2 * (gm0 + u)) = make_double2(e.x, -e.y);
            else if (e.x == 1.2345) sink[1] = e.y;
        }
        if (F & 4) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&full[slot])) : "memory");
    };
    if (F & 16) {
        for (int p = 0; p < 2; ++p) { if ((F & 8) && tid == 0) tma(p); gen_compute(p); gen_commit(p); }
    } else if (F & 4) {
        for (int p = 0; p < 2; ++p) { if ((F & 8) && tid == 0) tma(p); gen_store(p); }
    }
    for (int s = 0; s < stages; ++s) {
        const int slot = s % STAGES;
        if (F & 2) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                ge[0][u] = __ldg(tab + ((size_t)(s * 8 + gm0 + u) % 1000) * 512 + (grow & 511));
                ge[1][u] = __ldg(tab + 512000 + ((size_t)(s * 8 + gm0 + u) % 1000) * 512 + ((grow * 3) & 511));
            }
        }
        if ((F & 16) && s + 2 < stages) gen_compute(s + 2);
        if (F & 4) {
            mbar_wait(&full[slot], (s / STAGES) & 1);
            if ((F & 8) && tid == 0 && s + 2 < stages) tma(s + 2);
        }
        const double* A = sm + slot * STAGE_D;
        const double* B = A + TM * AST;
#pragma unroll
        for (int k4 = 0; k4 < KC / 2; ++k4) {
            double af[4], bf[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) af[i] = A[a_off + i * 8 * AST + 4 * k4];
#pragma unroll
            for (int j = 0; j < 8; ++j) bf[j] = B[b_off + 4 * k4 * BST + j * 8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
#pragma unroll
                for (int i = 0; i < 4; ++i) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
        if (F & 16) { if (s + 2 < stages) gen_commit(s + 2); }
        else if (s + 2 < stages) gen_store(s + 2);
    }
    double r = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) r += acc[i][j][0] + acc[i][j][1];
    if (r == 1.2345) sink[0] = r;
}

template <int F>
int run(const char* name, double* sink, int sms, const double2* tab, const double* btab)
{
    const int stages = 2000;
    const size_t smem = STAGES * STAGE_D * sizeof(double) + 64;
    CK(cudaFuncSetAttribute(loop<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        loop<F><<<sms, 256, smem>>>(sink, stages, tab, btab);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms;
    }
    CK(cudaGetLastError());
    const double fmas = (double)sms * 8 * stages * (KC / 2) * 32 * 256.0;
    printf("%-52s %.3f TFMA/s (%.1f%%)\n", name, fmas / (best * 1e-3) / 1e12,
           100.0 * fmas / (best * 1e-3) / (148.0 * 64 * 1.965e9));
    return 0;
}

int main()
{
    double* sink; CK(cudaMalloc(&sink, 64));
    double2* tab; CK(cudaMalloc(&tab, sizeof(double2) * 1024000)); CK(cudaMemset(tab, 0, sizeof(double2) * 1024000));
    double* btab; CK(cudaMalloc(&btab, sizeof(double) * 2000 * 512)); CK(cudaMemset(btab, 0, sizeof(double) * 2000 * 512));
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int n = p.multiProcessorCount;
    run<0>("DMMA + LDS only", sink, n, tab, btab);
    run<1>("+ FP64 ops", sink, n, tab, btab);
    run<2>("+ LDG", sink, n, tab, btab);
    run<4>("+ STS + mbarrier", sink, n, tab, btab);
    run<5>("+ FP64 + STS + mbarrier", sink, n, tab, btab);
    run<7>("+ FP64 + LDG + STS + mbarrier", sink, n, tab, btab);
    run<12>("+ STS + mbarrier + TMA", sink, n, tab, btab);
    run<15>("everything (FP64 + LDG + STS + mbarrier + TMA)", sink, n, tab, btab);
    run<20>("split phase: FP64 at stage top, STS at stage end", sink, n, tab, btab);
    return 0;
}
