// gsb_common.cuh -- error plumbing and sm_100a PTX helpers shared by the kernels.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdio>
#include <cstring>
#include <string>

#include "../../include/gsb200.h"

namespace gsb {

// ---------------------------------------------------------------------------------------------
// error handling: thread-local message, status codes, no exceptions across the C ABI
// ---------------------------------------------------------------------------------------------
inline std::string &last_error_ref()
{
    static thread_local std::string msg;
    return msg;
}

inline int fail(int code, const std::string &msg)
{
    last_error_ref() = msg;
    return code;
}

#define GSB_CUDA(call)                                                                         \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess) {                                                              \
            char buf__[512];                                                                   \
            snprintf(buf__, sizeof buf__, "%s failed: %s (%s:%d)", #call,                      \
                     cudaGetErrorString(e__), __FILE__, __LINE__);                             \
            return ::gsb::fail(GSB_ERR_CUDA, buf__);                                           \
        }                                                                                      \
    } while (0)

#define GSB_TRY(expr)                                                                          \
    do {                                                                                       \
        int rc__ = (expr);                                                                     \
        if (rc__ != GSB_OK) return rc__;                                                       \
    } while (0)

extern std::atomic<int64_t> g_launches;  // kernels launched by this library (gsb_get_counter)

// Function attributes (dynamic shared-memory size, carve-out) are per device: set them the first time a
// kernel family is launched on each device of the process, not once per process.
inline bool first_launch_on_device(std::atomic<uint64_t> &seen)
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return true;
    const uint64_t bit = 1ull << (dev & 63);
    return (seen.fetch_or(bit) & bit) == 0;
}

// ---------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + bulk async copy (TMA unit, SASS UBLKCP) on sm_100a
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)
                 : "memory");
}

// make barrier initialisation visible to the async (TMA) proxy
__device__ __forceinline__ void fence_barrier_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}

// add pending transaction bytes WITHOUT arriving (the lane arrives later, after its own stores)
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// 1-D bulk async copy global -> shared, completion signalled on an mbarrier (complete_tx).
// bytes must be a multiple of 16; src and dst 16-byte aligned.
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes,
                                         uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// ---------------------------------------------------------------------------------------------
// fused caller epilogue (gsb_epilogue in the C ABI): v = scale*sum, then + add[0][c], + add[1][c] ...
// Separately rounded multiply and adds (no FMA contraction), so the result has the same bits as the
// reference's numpy passes (generator.py:269-270, normalizer/tools.py:99-103) applied to the raw sum.
// ---------------------------------------------------------------------------------------------
struct Epi {
    double scale;
    double add[GSB_EPI_MAX_ADD][GSB_EPI_MAX_COMP];
    int n_add;
    int on;        // 0: store the raw sum
    // per-point step (gsb_point_epilogue; CondSRF.__call__, cond_srf.py:145-150), applied after the terms above:
    //   v = gain[i]*v;  v = offset[i] + v;  v = v + post[0]; ...     i = index of the point in its field
    const double *gain;
    const double *offset;
    double post[GSB_EPI_MAX_ADD];
    int n_post;
    int pp;        // 0: no per-point step
};

__device__ __forceinline__ double epi_apply(const Epi &e, double v, int comp, int64_t idx)
{
    if (!e.on) return v;
    v = __dmul_rn(e.scale, v);
    for (int k = 0; k < e.n_add; ++k) v = __dadd_rn(v, e.add[k][comp]);
    if (e.pp) {
        if (e.gain) v = __dmul_rn(__ldg(e.gain + idx), v);
        if (e.offset) v = __dadd_rn(__ldg(e.offset + idx), v);
        for (int k = 0; k < e.n_post; ++k) v = __dadd_rn(v, e.post[k]);
    }
    return v;
}

inline Epi make_epi(const gsb_epilogue *src, const gsb_point_epilogue *pp = nullptr)
{
    Epi e;
    std::memset(&e, 0, sizeof e);
    if (!src && !pp) return e;
    e.on = 1;
    e.scale = 1.0;
    if (src) {
        e.scale = src->scale;
        e.n_add = src->n_add;
        for (int k = 0; k < GSB_EPI_MAX_ADD; ++k)
            for (int c = 0; c < GSB_EPI_MAX_COMP; ++c) e.add[k][c] = src->add[k][c];
    }
    if (pp) {
        e.pp = 1;
        e.gain = pp->gain;
        e.offset = pp->offset;
        e.n_post = pp->n_add;
        for (int k = 0; k < GSB_EPI_MAX_ADD; ++k) e.post[k] = pp->add[k];
    }
    return e;
}

// the same epilogue for the points [first, ...) of the field (host route: point chunks)
inline Epi epi_shift(Epi e, int64_t first)
{
    if (e.gain) e.gain += first;
    if (e.offset) e.offset += first;
    return e;
}

__device__ __forceinline__ int hi32(double x) { return __double2hiint(x); }
__device__ __forceinline__ int lo32(double x) { return __double2loint(x); }

}  // namespace gsb
