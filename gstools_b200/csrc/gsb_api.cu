// gsb_api.cu -- C ABI (include/gsb200.h) over the sm_100a kernels.
//
// Host side of the drop-in boundary: argument validation, staging of host buffers, scratch
// memory from the stream-ordered pool, kernel selection.  No compute happens on the CPU and
// there is no fallback: without a CUDA device every compute entry returns GSB_ERR_NO_DEVICE.
#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include <cuda.h>

#include "gsb_common.cuh"
#include "gsb_direct.cuh"
#include "gsb_separable.cuh"
#include "gsb_sepk.cuh"
#include "gsb_krige.cuh"
#include "gsb_sampler.cuh"

namespace gsb {

std::atomic<int64_t> g_launches{0};
static std::atomic<int64_t> g_opt_structured_min_tiles{0};   // meshes with fewer 128 x 128 tiles go to the direct kernel
static std::atomic<int64_t> g_opt_force_path{0};
static std::atomic<int64_t> g_opt_host_chunk_points{1 << 22};
static std::atomic<int64_t> g_opt_scratch_mb{3072};  // scratch budget of the kriging right-hand sides (two buffers)
static std::atomic<int64_t> g_opt_fold_axes{1};     // 0 never, 1 when it improves the tile utilisation, 2 whenever it fits
static std::atomic<int64_t> g_cnt_folded{0};
static std::atomic<int64_t> g_opt_direct_cfg{-1};   // -1 auto, else force P = 1 / 2 / 8 points per thread (0 / 1 / 2)
static std::atomic<int64_t> g_opt_direct_split{1};  // 0: never split the mode loop of small point sets over CTAs
static std::atomic<int64_t> g_cnt_direct{0}, g_cnt_separable{0};
// optional device-side timing of the dominant kernels (bench.py roofline): events recorded on the
// launch stream around every direct / separable launch while the option "time_kernels" is 1
static std::atomic<int64_t> g_opt_time_kernels{0};
static std::mutex g_time_mutex;
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_time_events;

// debugging aid (option "trace" = 1): start/end of every agen / contract kernel relative to the
// start of the structured call, printed to stderr after a device synchronisation
static std::atomic<int64_t> g_opt_trace{0};
struct TraceRec { const char *name; cudaEvent_t e0, e1; };
static std::vector<TraceRec> g_trace;
struct TraceScope {
    cudaEvent_t e0 = nullptr, e1 = nullptr; cudaStream_t st; const char *name;
    TraceScope(const char *n, cudaStream_t s) : st(s), name(n)
    {
        if (!g_opt_trace.load()) return;
        cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, st);
    }
    ~TraceScope() { if (e0) { cudaEventRecord(e1, st); g_trace.push_back({name, e0, e1}); } }
};

struct KernelTimer {
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    cudaStream_t st;
    explicit KernelTimer(cudaStream_t s) : st(s)
    {
        if (!g_opt_time_kernels.load()) return;
        if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) { e0 = e1 = nullptr; return; }
        cudaEventRecord(e0, st);
    }
    ~KernelTimer()
    {
        if (!e0) return;
        cudaEventRecord(e1, st);
        std::lock_guard<std::mutex> lock(g_time_mutex);
        g_time_events.emplace_back(e0, e1);
    }
};

// ---------------------------------------------------------------------------------------------
// device bookkeeping
// ---------------------------------------------------------------------------------------------
struct DeviceState {
    static constexpr int N_STREAMS = 4;        // 0: host-path main, 1: copies, 2/3: contraction
    static constexpr int N_EVENTS = 5;
    static constexpr int N_CHUNK_EVENTS = 8;
    bool ready = false;
    int sm_count = 0;
    cudaStream_t streams[N_STREAMS] = {};
    cudaEvent_t events[N_EVENTS] = {};
    cudaEvent_t chunk_events[N_CHUNK_EVENTS] = {};      // A generation of chunk c done
    cudaEvent_t contract_events[N_CHUNK_EVENTS] = {};   // contraction of chunk c done
    std::mutex call_mutex;                     // one structured / host-route call at a time per device
};
static std::mutex g_dev_mutex;
static DeviceState g_dev[64];

// restores the caller's current device when an API call returns (torch tracks it too)
struct DeviceGuard {
    int prev = -1;
    DeviceGuard() { if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; } }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

static int ensure_device(int device, DeviceState **out)
{
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        return fail(GSB_ERR_NO_DEVICE,
                    "no CUDA device available: the B200 backend has no CPU fallback");
    }
    if (device < 0 || device >= count || device >= 64)
        return fail(GSB_ERR_ARGUMENT, "invalid device index");
    GSB_CUDA(cudaSetDevice(device));
    std::lock_guard<std::mutex> lock(g_dev_mutex);
    DeviceState &d = g_dev[device];
    if (!d.ready) {
        cudaDeviceProp prop;
        GSB_CUDA(cudaGetDeviceProperties(&prop, device));
        if (prop.major < 10)
            return fail(GSB_ERR_NO_DEVICE, std::string("device '") + prop.name +
                                               "' is not sm_100-class; this library is built for sm_100a only");
        d.sm_count = prop.multiProcessorCount;
        // Stream 2 runs the A generation.  It gets the highest priority: the block scheduler
        // finishes dispatching one grid before it starts the next of the same priority, so without
        // this the A generation of chunk c+1 would only start in the tail of contraction c.
        int prio_lo = 0, prio_hi = 0;
        GSB_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        for (int i = 0; i < DeviceState::N_STREAMS; ++i)
            GSB_CUDA(cudaStreamCreateWithPriority(&d.streams[i], cudaStreamNonBlocking, i == 2 ? prio_hi : prio_lo));
        for (int i = 0; i < DeviceState::N_EVENTS; ++i)
            GSB_CUDA(cudaEventCreateWithFlags(&d.events[i], cudaEventDisableTiming));
        for (int i = 0; i < DeviceState::N_CHUNK_EVENTS; ++i) {
            GSB_CUDA(cudaEventCreateWithFlags(&d.chunk_events[i], cudaEventDisableTiming));
            GSB_CUDA(cudaEventCreateWithFlags(&d.contract_events[i], cudaEventDisableTiming));
        }
        // keep freed scratch in the stream-ordered pool instead of returning it to the driver
        cudaMemPool_t pool;
        GSB_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
        uint64_t thresh = UINT64_MAX;
        GSB_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thresh));
        d.ready = true;
    }
    *out = &d;
    return GSB_OK;
}

// scratch allocation on a stream (freed on the same stream when the holder dies)
struct Scratch {
    cudaStream_t st;
    std::vector<void *> ptrs;
    explicit Scratch(cudaStream_t s) : st(s) {}
    ~Scratch()
    {
        for (void *p : ptrs) cudaFreeAsync(p, st);
    }
    template <typename T>
    int alloc(T **p, size_t count)
    {
        void *q = nullptr;
        GSB_CUDA(cudaMallocAsync(&q, std::max<size_t>(count, 1) * sizeof(T), st));
        ptrs.push_back(q);
        *p = static_cast<T *>(q);
        return GSB_OK;
    }
};

static int check_epilogue(const gsb_epilogue *e)
{
    if (e && (e->n_add < 0 || e->n_add > GSB_EPI_MAX_ADD))
        return fail(GSB_ERR_ARGUMENT, "epilogue: n_add must be in 0..GSB_EPI_MAX_ADD");
    return GSB_OK;
}

static int check_point_epilogue(const gsb_point_epilogue *e, bool vec)
{
    if (!e) return GSB_OK;
    if (vec) return fail(GSB_ERR_ARGUMENT, "point epilogue: scalar fields only (cond_srf.py:56)");
    if (e->n_add < 0 || e->n_add > GSB_EPI_MAX_ADD)
        return fail(GSB_ERR_ARGUMENT, "point epilogue: n_add must be in 0..GSB_EPI_MAX_ADD");
    return GSB_OK;
}

static int check_common(const void *cov, const void *z1, const void *z2, int dim, int64_t n_modes)
{
    if (dim < 1 || dim > GSB_MAX_DIM) return fail(GSB_ERR_ARGUMENT, "dim must be in 1..8");
    if (n_modes < 0) return fail(GSB_ERR_ARGUMENT, "n_modes must be >= 0");
    if (n_modes > 0 && (!cov || !z1 || !z2))
        return fail(GSB_ERR_ARGUMENT, "cov_samples, z_1 and z_2 must not be NULL");
    return GSB_OK;
}

// ---------------------------------------------------------------------------------------------
// direct path on device-resident data
// ---------------------------------------------------------------------------------------------
static int pack_modes(const double *d_cov, const double *d_z1, const double *d_z2, const double *d_sf,
                      int dim, int64_t n_modes, bool vec, double **d_recs, int64_t *n_modes_pad,
                      Scratch &scr, cudaStream_t st)
{
    const int64_t pad = std::max<int64_t>(4, (n_modes + 3) / 4 * 4);
    GSB_TRY(scr.alloc(d_recs, (size_t)pad * direct_rec(dim, vec)));
    const int threads = 128;
    const int blocks = (int)std::min<int64_t>((pad + threads - 1) / threads, 1024);
    pack_modes_kernel<<<blocks, threads, 0, st>>>(d_cov, d_z1, d_z2, d_sf, dim, n_modes, pad, vec ? 1 : 0,
                                                  *d_recs);
    g_launches.fetch_add(1);
    GSB_CUDA(cudaGetLastError());
    *n_modes_pad = pad;
    return GSB_OK;
}

// d_recs already packed; evaluates n_pts points
static int direct_on_device(const double *d_recs, int64_t n_modes_pad, const double *d_pos,
                            int64_t pos_ld, int dim, bool vec, int64_t n_pts, double *d_out,
                            int64_t out_ld, const Epi &epi, const DeviceState &dev, Scratch &scr,
                            cudaStream_t st, bool allow_tail_split = true)
{
    if (n_pts == 0) return GSB_OK;
    // Launch configuration by a small cost model: time ~ waves x points per SM and wave / relative speed.  The
    // big-CTA configuration is the fastest per point (measured 1.02 vs 0.93 vs 0.77 Tpair/s, 2-D) but
    // quantises into waves of 2 x 1024 points per SM; mid-sized point sets are better off with smaller CTAs.
    // (A point's bits do not depend on the configuration: the per-point instruction sequence is the same.)
    const int64_t want = 2LL * dev.sm_count;
    static const double speed[3] = {0.756, 0.906, 1.00};   // tools/direct_size_sweep.py
    static const int resident[3] = {8, 4, 2};
    auto model_time = [&](int c, int64_t n) -> double {
        const int64_t ppc = direct_cfg_points(c, dim);
        const int64_t nc = (n + ppc - 1) / ppc;
        const int64_t slots = (int64_t)resident[c] * dev.sm_count;
        return (double)((nc + slots - 1) / slots) * (double)(resident[c] * ppc) / speed[c];
    };
    auto best_cfg = [&](int64_t n, double *t_out) -> int {
        double best = 1e300;
        int cfg = 0;
        for (int c = 0; c < 3; ++c) {
            const double t = model_time(c, n);
            if (t < best * 0.999) {
                best = t;
                cfg = c;
            }
        }
        if (t_out) *t_out = best;
        return cfg;
    };
    double t_single = 0.0;
    int cfg = best_cfg(n_pts, &t_single);
    if (g_opt_direct_cfg.load() >= 0) cfg = (int)g_opt_direct_cfg.load();
    // Tail split: a point set of 8.25 waves of the big configuration would run 9 (2.5 M points per rank of config 3
    // on eight GPUs: 92 % of the time useful).  The full waves keep that configuration; the remaining quarter wave goes
    // to a second launch whose smaller CTAs spread it over all SMs.  Same bits: the per-point sequence does not change.
    if (allow_tail_split && g_opt_direct_cfg.load() < 0) {
        const int64_t ppc = direct_cfg_points(cfg, dim);
        const int64_t wave_pts = (int64_t)resident[cfg] * dev.sm_count * ppc;
        const int64_t n_main = n_pts / wave_pts * wave_pts;
        if (n_main > 0 && n_main < n_pts) {
            double t_tail = 0.0;
            best_cfg(n_pts - n_main, &t_tail);
            const double t_split = model_time(cfg, n_main) + t_tail + 0.02 * model_time(cfg, wave_pts);   // + one launch
            if (t_split < 0.98 * t_single) {
                GSB_TRY(direct_on_device(d_recs, n_modes_pad, d_pos, pos_ld, dim, vec, n_main, d_out, out_ld, epi, dev, scr,
                                         st, false));
                GSB_TRY(direct_on_device(d_recs, n_modes_pad, d_pos + n_main, pos_ld, dim, vec, n_pts - n_main,
                                         d_out + n_main, out_ld, epi_shift(epi, n_main), dev, scr, st, false));
                g_cnt_direct.fetch_sub(1);      // one call, two launches
                return GSB_OK;
            }
        }
    }
    const int64_t ctas = (n_pts + direct_cfg_points(cfg, dim) - 1) / direct_cfg_points(cfg, dim);
    const int n_tiles = (int)((n_modes_pad + DIRECT_TM - 1) / DIRECT_TM);
    int n_split = 1;
    // Mode splitting for small point sets (config 1: 10 000 points leave seven eighths of the thread slots empty and
    // every thread walks 1000 modes alone): the mode tiles are dealt out to gridDim.y CTAs per point block, each stores
    // its tile sums, and the reduce kernel adds them in the unsplit kernel's order -- same bits, ~4x less latency.
    const int64_t thread_slots = 8LL * 64 * dev.sm_count;
    const int ncomp_split = vec ? dim : 1;
    if (cfg == 0 && n_tiles >= 2 && 2 * n_pts < thread_slots && g_opt_direct_split.load() != 0 &&
        (int64_t)n_tiles * ncomp_split * n_pts <= ((int64_t)1 << 25))          // tile sums: at most 256 MiB of scratch
        n_split = (int)std::min<int64_t>(n_tiles, (thread_slots + n_pts - 1) / n_pts);
    (void)want;
    DirectParams prm;
    prm.recs = d_recs;
    prm.n_modes_pad = n_modes_pad;
    prm.pos = d_pos;
    prm.pos_ld = pos_ld;
    prm.n_pts = n_pts;
    prm.out = d_out;
    prm.out_ld = out_ld;
    prm.n_split = n_split;
    prm.partial = nullptr;
    prm.epi = epi;
    const int ncomp = vec ? dim : 1;
    if (n_split > 1) GSB_TRY(scr.alloc(&prm.partial, (size_t)n_tiles * ncomp * n_pts));
    {
        KernelTimer timer(st);
        GSB_TRY(launch_direct(dim, vec, prm, cfg, st));
    }
    if (n_split > 1) {
        dim3 grid((unsigned)((n_pts + 255) / 256), (unsigned)ncomp);
        reduce_partials_kernel<<<grid, 256, 0, st>>>(prm.partial, n_tiles, ncomp, n_pts, d_out, out_ld, epi);
        g_launches.fetch_add(1);
        GSB_CUDA(cudaGetLastError());
    }
    g_cnt_direct.fetch_add(1);
    return GSB_OK;
}

static int summate_impl(const double *cov, const double *z1, const double *z2, const double *sf,
                        const double *pos, int64_t pos_ld, int dim, int64_t n_modes, int64_t n_pts,
                        double *out, int64_t out_ld, bool vec, const gsb_epilogue *epilogue, int mem,
                        int device, void *stream, const gsb_point_epilogue *pepi = nullptr, int64_t epi_first = 0)
{
    DeviceGuard guard;
    GSB_TRY(check_common(cov, z1, z2, dim, n_modes));
    GSB_TRY(check_epilogue(epilogue));
    GSB_TRY(check_point_epilogue(pepi, vec));
    // epi_first: the call evaluates the points [epi_first, epi_first + n_pts) of a bigger field (multi-GPU plan);
    // the per-point arrays of `pepi` describe the whole field
    const Epi epi = epi_shift(make_epi(epilogue, pepi), epi_first);
    if (n_pts < 0) return fail(GSB_ERR_ARGUMENT, "n_pts must be >= 0");
    if (vec && dim != 2 && dim != 3)
        return fail(GSB_ERR_ARGUMENT,
                    "summate_incompr: dim must be 2 or 3 (generator.py:514-517)");
    if (n_pts > 0 && (!pos || !out)) return fail(GSB_ERR_ARGUMENT, "pos and out must not be NULL");
    if (pos_ld < n_pts || (vec && out_ld < n_pts))
        return fail(GSB_ERR_ARGUMENT, "leading dimension smaller than n_pts");
    if (mem != GSB_MEM_HOST && mem != GSB_MEM_DEVICE)
        return fail(GSB_ERR_ARGUMENT, "mem must be GSB_MEM_HOST or GSB_MEM_DEVICE");
    if (n_pts == 0) return GSB_OK;
    DeviceState *dev = nullptr;
    GSB_TRY(ensure_device(device, &dev));
    const int ncomp = vec ? dim : 1;

    if (mem == GSB_MEM_DEVICE) {
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        Scratch scr(st);
        double *d_recs = nullptr;
        int64_t pad = 0;
        GSB_TRY(pack_modes(cov, z1, z2, sf, dim, n_modes, vec, &d_recs, &pad, scr, st));
        return direct_on_device(d_recs, pad, pos, pos_ld, dim, vec, n_pts, out, out_ld, epi, *dev, scr, st);
    }

    // ---- host buffers: stage modes once, then pipeline point chunks over two streams ----
    // (the library's own streams and events are per device: one host-route call at a time)
    std::lock_guard<std::mutex> lock(dev->call_mutex);
    cudaStream_t s0 = dev->streams[0];
    Scratch scr0(s0);
    double *d_cov, *d_z1, *d_z2, *d_recs, *d_sf = nullptr;
    GSB_TRY(scr0.alloc(&d_cov, (size_t)dim * n_modes));
    GSB_TRY(scr0.alloc(&d_z1, (size_t)n_modes));
    GSB_TRY(scr0.alloc(&d_z2, (size_t)n_modes));
    if (sf) {
        GSB_TRY(scr0.alloc(&d_sf, (size_t)n_modes));
        if (n_modes > 0)
            GSB_CUDA(cudaMemcpyAsync(d_sf, sf, sizeof(double) * n_modes, cudaMemcpyHostToDevice, s0));
    }
    if (n_modes > 0) {
        GSB_CUDA(cudaMemcpyAsync(d_cov, cov, sizeof(double) * dim * n_modes, cudaMemcpyHostToDevice, s0));
        GSB_CUDA(cudaMemcpyAsync(d_z1, z1, sizeof(double) * n_modes, cudaMemcpyHostToDevice, s0));
        GSB_CUDA(cudaMemcpyAsync(d_z2, z2, sizeof(double) * n_modes, cudaMemcpyHostToDevice, s0));
    }
    int64_t pad = 0;
    GSB_TRY(pack_modes(d_cov, d_z1, d_z2, d_sf, dim, n_modes, vec, &d_recs, &pad, scr0, s0));
    GSB_CUDA(cudaEventRecord(dev->events[0], s0));

    const int64_t chunk = std::min<int64_t>(n_pts, std::max<int64_t>(1024, g_opt_host_chunk_points.load()));
    const int nbuf = (n_pts > chunk) ? 2 : 1;
    Scratch scr1(dev->streams[1]);
    double *d_pos[2] = {nullptr, nullptr}, *d_out[2] = {nullptr, nullptr};
    for (int b = 0; b < nbuf; ++b) {
        Scratch &s = b ? scr1 : scr0;
        GSB_TRY(s.alloc(&d_pos[b], (size_t)dim * chunk));
        GSB_TRY(s.alloc(&d_out[b], (size_t)ncomp * chunk));
    }
    GSB_CUDA(cudaStreamWaitEvent(dev->streams[1], dev->events[0], 0));
    int64_t c = 0;
    for (int64_t i0 = 0; i0 < n_pts; i0 += chunk, ++c) {
        const int b = (int)(c % nbuf);
        cudaStream_t st = dev->streams[b];
        Scratch &s = b ? scr1 : scr0;
        const int64_t m = std::min(chunk, n_pts - i0);
        GSB_CUDA(cudaMemcpy2DAsync(d_pos[b], sizeof(double) * chunk, pos + i0, sizeof(double) * pos_ld,
                                   sizeof(double) * m, dim, cudaMemcpyHostToDevice, st));
        GSB_TRY(direct_on_device(d_recs, pad, d_pos[b], chunk, dim, vec, m, d_out[b], chunk, epi_shift(epi, i0), *dev,
                                 s, st));
        GSB_CUDA(cudaMemcpy2DAsync(out + i0, sizeof(double) * (vec ? out_ld : n_pts), d_out[b],
                                   sizeof(double) * chunk, sizeof(double) * m, ncomp,
                                   cudaMemcpyDeviceToHost, st));
    }
    GSB_CUDA(cudaStreamSynchronize(dev->streams[1]));
    GSB_CUDA(cudaStreamSynchronize(s0));
    return GSB_OK;
}

// ---------------------------------------------------------------------------------------------
// structured path
// ---------------------------------------------------------------------------------------------
struct SlabOpt { int64_t lo, hi; };     // range of axis 0 evaluated by one device of a multi-GPU plan

struct MeshInfo {
    int dim;
    int64_t len[GSB_MAX_DIM];
    int64_t off[GSB_MAX_DIM];
    int64_t total_axes;  // sum(len)
    int64_t n;           // prod(len)
    int64_t n_rows;      // prod(len[:-1])
    double matrix[GSB_MAX_DIM * GSB_MAX_DIM];
    bool identity;
};

// ---- second-generation separable path (gsb_sepk.cuh): tables -> ONE stream-K contraction launch per piece ----
static std::atomic<int64_t> g_cnt_sk{0};
static std::atomic<int64_t> g_opt_sk_table_mb{256};     // cap on the tile-axis table (all batch entries)
static std::atomic<int64_t> g_opt_sk_grid{0};           // 0: one CTA per SM; else a fixed grid (tests)
static std::atomic<int64_t> g_opt_sk_pack{0};           // row packing across slow indices: 0 auto, 1 never, 2 whenever possible
static std::atomic<int64_t> g_cnt_packed{0};
static std::atomic<int64_t> g_opt_host_pieces{0};       // host route: 0 = growing pieces (default), n = n equal pieces

// The virtual mesh of the contraction: [outer slow axes | tile axes (folded: ly rows) | inner slow axes |
// column axes (folded: lc columns)]
struct SkLayout {
    int n_prefix, n_tile_axes, n_inner, n_col_axes;
    int64_t n_slow, n_in, ly, lc;
    int n_ytiles, n_col_tiles;
    // virtual row space of a field (gsb_sepk.cuh): lyp rows per slow index, n_vt row tiles per field
    bool pack;
    int64_t lyp, n_vt;
};

// Cost of a field's row tiles when they are packed across the slow indices: all tiles full except the last one.
static int64_t sk_packed_cost(int64_t n_slow, int64_t ly, int64_t lc, int n_col_tiles, int64_t *n_vt_out)
{
    const int64_t lyp = (ly + 7) / 8 * 8, vrows = n_slow * lyp, n_vt = (vrows + SK_TM - 1) / SK_TM;
    int64_t per = 0;
    for (int ct = 0; ct < n_col_tiles; ++ct)
        per += (n_vt - 1) * sk_tile_cost(vrows, lc, 0, ct) + sk_tile_cost(vrows, lc, (int)(n_vt - 1), ct);
    if (n_vt_out) *n_vt_out = n_vt;
    return per;
}

// Which contiguous group of row axes becomes the tile axis?  Estimated time = contraction (cost units of sk_plan:
// 64 units = one full 128 x 128 tile stage = 4096 SM cycles at 64 DFMA / clk) + building the tile-axis table
// (one sincos and 24 bytes per entry).  E.g. 100^3: the middle axis alone (78 % of a tile) beats folding both row
// axes (99 %, but a 237 MB table for a 8 MB field); 512 x 16 x 512: axis 0 alone; 40 x 50 x 130: both folded.
static SkLayout sk_choose_layout(const MeshInfo &mesh, int64_t n_modes_pad, int64_t n_batch, int ncomp, bool fold_cols,
                                 int sm_count, bool host_route, double *est_seconds = nullptr)
{
    const int dim = mesh.dim;
    SkLayout L;
    L.n_col_axes = fold_cols ? 2 : 1;
    L.lc = fold_cols ? mesh.len[dim - 2] * mesh.len[dim - 1] : mesh.len[dim - 1];
    L.n_col_tiles = (int)((L.lc + SK_TN - 1) / SK_TN);
    const int n_row = dim - L.n_col_axes;                   // >= 1
    const int64_t n_stages = n_modes_pad / SK_KC;
    const int64_t cap = g_opt_sk_table_mb.load() << 20;
    int64_t n_rows = 1;
    for (int t = 0; t < n_row; ++t) n_rows *= mesh.len[t];
    double best = 1e300;
    int best_a = n_row - 1, best_b = n_row;
    bool best_pack = false;
    const int64_t pack_opt = g_opt_sk_pack.load();
    for (int a = 0; a < n_row; ++a) {
        int64_t P = 1;
        for (int b = a + 1; b <= n_row; ++b) {
            P *= mesh.len[b - 1];
            const int64_t n_yt = (P + SK_TM - 1) / SK_TM;
            const int64_t bytes = n_batch * n_yt * n_stages * SK_A_TILE * (int64_t)sizeof(double);
            const bool single = (b == a + 1);
            if (n_yt > (1 << 20) || (!single && bytes > cap)) break;
            // cost of one (field, slow index) period: full row tiles + the last one, over all column tiles
            int64_t per = 0;
            for (int ct = 0; ct < L.n_col_tiles; ++ct)
                per += (n_yt - 1) * sk_tile_cost(P, L.lc, 0, ct) + sk_tile_cost(P, L.lc, (int)(n_yt - 1), ct);
            double units = (double)(n_rows / P) * (double)per * (double)n_stages * (double)(n_batch * ncomp);
            const bool no_slow = (b - a == n_row);
            // Row tiles packed across the slow indices (200^3: 200 rows per plane instead of 256): possible when
            // slow axes exist (the per-slow factors are what a tile then has several of), P is not a multiple of 128
            // anyway, and a 128-row tile touches at most SK_MAXSEG slow indices
            bool pack = false;
            if (!no_slow && n_rows / P > 1 && P % SK_TM != 0 && (P + 7) / 8 * 8 >= SK_PACK_MIN_ROWS && pack_opt != 1) {
                int64_t n_vt = 0;
                const int64_t pper = sk_packed_cost(n_rows / P, P, L.lc, L.n_col_tiles, &n_vt);
                const double punits = (double)pper * (double)n_stages * (double)(n_batch * ncomp);
                // (a packed tile costs a little more than a plain one -- one bulk copy and one block of factors per
                // slow index it touches, one more LDS per quad of modes: 64 x 512 x 512 with axis 0 packed measured
                // 4 % slower than axis 1 as plain tile axis, which the table sizes alone would not have chosen)
                const double penalty = 1.05;
                if (n_vt * L.n_col_tiles < (1 << 26) && (punits * penalty < units || pack_opt == 2)) {
                    pack = true;
                    units = punits * penalty;
                }
            }
            // (a tile is cut into at most ~8 shares, sk_plan: a tiny mesh cannot use every SM)
            const double min_share = std::max<double>(64.0, (double)n_stages * 64.0 / 8.0);
            const double ctas = std::max(1.0, std::min((double)sm_count, std::floor(units / min_share)));
            const double t_contract = units * 64.0 / (ctas * 1.9e9 * (no_slow ? 0.96 : 0.93));
            const double t_table = (double)bytes / 1.0e12;
            // Host route: the result leaves in contiguous pieces while the next piece contracts (sk_on_device).  With
            // slow axes INSIDE the tile axis a (field, slow index) unit is not contiguous in the output, pieces are
            // whole fields, and the copy of the last field (PCIe, ~50 GB/s) is exposed instead of ~1/10 of it.
            // (256 x 512 x 512 on one device of a 2-GPU plan: axis 0 as tile axis has the smaller table, but
            // its ONE piece cost 7.8 ms + 9.4 ms instead of 10 ms overlapped.)
            double t_copy = 0.0;
            if (host_route) {
                const double field_bytes = (double)n_rows * (double)L.lc * sizeof(double);
                t_copy = field_bytes / 50e9 * ((b < n_row) ? 1.0 : 0.1);
            }
            // ties: prefer the trailing axes (rows of a tile then are neighbours in memory)
            const double t = (t_contract + t_table + t_copy) * (1.0 + 1e-6 * (n_row - b));
            if (t < best) { best = t; best_a = a; best_b = b; best_pack = pack; }
        }
    }
    L.n_prefix = best_a;
    L.n_tile_axes = best_b - best_a;
    L.n_inner = n_row - best_b;
    L.ly = 1;
    for (int t = best_a; t < best_b; ++t) L.ly *= mesh.len[t];
    L.n_in = 1;
    for (int t = best_b; t < n_row; ++t) L.n_in *= mesh.len[t];
    L.n_slow = n_rows / L.ly;
    L.n_ytiles = (int)((L.ly + SK_TM - 1) / SK_TM);
    L.pack = best_pack;
    L.lyp = best_pack ? (L.ly + 7) / 8 * 8 : (int64_t)L.n_ytiles * SK_TM;
    L.n_vt = (L.n_slow * L.lyp + SK_TM - 1) / SK_TM;
    if (est_seconds) {
        const double b_bytes = (double)n_batch * ncomp * L.n_col_tiles * (double)n_stages * SK_B_TILE * sizeof(double);
        *est_seconds = best + b_bytes / 1.0e12 + 8e-6;      // + two launches
    }
    return L;
}

static int sk_on_device(const double *d_cov, const double *d_z1, const double *d_z2, const double *d_sf,
                        const double *d_axes, const MeshInfo &mesh, const SkLayout &L, int64_t n_modes, int64_t n_batch,
                        bool vec, const Epi &epi, double *d_out, int64_t d_fstride, double *h_out, int64_t h_fstride,
                        DeviceState &dev, cudaStream_t st)
{
    const int dim = mesh.dim;
    const int ncomp = vec ? dim : 1;
    if (n_batch > 65535) return fail(GSB_ERR_ARGUMENT, "n_batch too large");
    Scratch scr(st);
    const int n_modes_pad = (int)((n_modes + SK_KC - 1) / SK_KC * SK_KC);
    const int n_stages = n_modes_pad / SK_KC;
    const bool scale = (L.n_prefix + L.n_inner) > 0;     // slow axes exist (even of length 1: their phase counts)

    const int64_t gopt = g_opt_sk_grid.load();
    const int max_grid = gopt > 0 ? (int)std::min<int64_t>(gopt, SK_MAX_GRID) : std::min(dev.sm_count, SK_MAX_GRID);
    // One launch for everything on the device route.  The host route cuts the (field, slow index) units into
    // pieces so that the D2H copy of a piece overlaps the contraction of the next.  PCIe is the slower side (the
    // 512^3 field: 15.6 ms of contraction, 20 ms of copy at 53 GB/s), so the first piece is small -- the copy
    // engine starts early -- and the pieces grow by 1.25x, the ratio at which the next contraction still finishes
    // before the previous copy does.  Unit u = z * n_slow + slow covers the output elements [u, u + 1) * ly * lc,
    // so a piece is ONE contiguous copy (with inner slow axes a unit is not contiguous: pieces then are whole fields).
    // Work is numbered in ROW TILES: field z = (batch, component) owns the row tiles [z * n_vt, (z + 1) * n_vt), each
    // of n_col_tiles tiles.  A piece is a range of row tiles [g0, g1); after it, the (field, slow index) units
    // [done(g0), done(g1)) are complete -- unit u covers the output elements [u, u + 1) * ly * lc, so a piece leaves
    // as ONE contiguous copy per field (with inner slow axes a unit is not contiguous: pieces then are whole fields).
    const int64_t n_z = n_batch * ncomp;
    const int64_t total_vt = n_z * L.n_vt;
    // cut granularity: whole fields with inner slow axes; whole slow indices when nothing is packed
    const int64_t quantum = L.n_in > 1 ? L.n_vt : (L.pack ? 1 : (int64_t)L.n_ytiles);
    auto done_units = [&](int64_t g) -> int64_t {
        const int64_t z = g / L.n_vt, vt = g % L.n_vt;
        if (L.n_in > 1) return z * L.n_slow;
        return z * L.n_slow + std::min<int64_t>(L.n_slow, vt * SK_TM / L.lyp);
    };
    std::vector<int64_t> cuts{0};      // piece k = row tiles [cuts[k], cuts[k+1])
    if (h_out) {
        const int64_t n_q = total_vt / quantum;                                       // cut positions available
        const int64_t tiles_per_q = quantum * L.n_col_tiles;
        const int64_t min_cut = std::max<int64_t>(1, ((int64_t)max_grid + tiles_per_q - 1) / tiles_per_q);   // >= one wave
        double want = std::max<double>((double)min_cut, (double)n_q / 64.0);
        const int64_t fixed = g_opt_host_pieces.load();
        for (int64_t k = 1; fixed > 0 && k <= fixed; ++k) cuts.push_back(n_q * k / fixed * quantum);
        while (fixed <= 0 && cuts.back() < total_vt && cuts.size() < 32) {
            cuts.push_back(std::min<int64_t>(total_vt, cuts.back() + std::max<int64_t>(min_cut, (int64_t)want) * quantum));
            want *= 1.25;
        }
    }
    if (cuts.size() == 1) cuts.push_back(total_vt);      // device route: one piece
    cuts.back() = total_vt;
    const int64_t pieces = (int64_t)cuts.size() - 1;
    SkTableParams tp;
    std::memset(&tp, 0, sizeof tp);
    tp.cov = d_cov; tp.z1 = d_z1; tp.z2 = d_z2; tp.sf = d_sf; tp.axes = d_axes;
    for (int t = 0; t < dim; ++t) {
        tp.axis_off[t] = mesh.off[t];
        tp.axis_len[t] = mesh.len[t];
    }
    std::memcpy(tp.matrix, mesh.matrix, sizeof tp.matrix);
    tp.dim = dim;
    tp.n_prefix = L.n_prefix;
    tp.n_tile_axes = L.n_tile_axes;
    tp.n_inner = L.n_inner;
    tp.n_col_axes = L.n_col_axes;
    tp.ly = L.ly;
    tp.lc = L.lc;
    tp.n_slow = L.n_slow;
    tp.n_in = L.n_in;
    tp.n_modes = n_modes;
    tp.n_modes_pad = n_modes_pad;
    tp.ncomp = ncomp;
    tp.n_ytiles = L.n_ytiles;
    tp.n_col_tiles = L.n_col_tiles;
    // one scratch block per call (every stream-ordered allocation costs host time, which is what a small mesh sees)
    auto pad16 = [](size_t doubles) { return (doubles + 15) / 16 * 16; };
    const size_t n_t = pad16((size_t)n_batch * L.n_ytiles * n_stages * SK_A_TILE);
    const size_t n_b = pad16((size_t)n_batch * ncomp * L.n_col_tiles * n_stages * SK_B_TILE);
    const size_t n_c = scale ? pad16((size_t)n_batch * L.n_slow * n_modes_pad * 2) : 0;
    const size_t n_s = pad16((size_t)max_grid * SK_TM * SK_TN);
    tp.n_flags = (int)(pieces * max_grid);
    const size_t n_f = pad16(((size_t)tp.n_flags + 1) / 2);
    double *block = nullptr;
    GSB_TRY(scr.alloc(&block, n_t + n_b + n_c + n_s + n_f));
    tp.ttab = block;
    tp.btile = block + n_t;
    tp.ctab = scale ? reinterpret_cast<double2 *>(block + n_t + n_b) : nullptr;
    double *slots = block + n_t + n_b + n_c;
    tp.flags = reinterpret_cast<unsigned *>(block + n_t + n_b + n_c + n_s);
    {
        const int64_t max_width = std::max<int64_t>(std::max<int64_t>((int64_t)L.n_ytiles * SK_TM,
                                                                       (int64_t)L.n_col_tiles * SK_TN),
                                                    scale ? L.n_slow : 1);
        const int64_t work = max_width * n_modes_pad;
        tp.small_index = (work < ((int64_t)1 << 31) && L.ly < ((int64_t)1 << 31) && L.lc < ((int64_t)1 << 31) &&
                          L.n_slow < ((int64_t)1 << 31)) ? 1 : 0;
        dim3 grid((unsigned)std::min<int64_t>((work + 255) / 256, 4096), scale ? 3u : 2u, (unsigned)n_batch);
        sk_tables_kernel<<<grid, 256, 0, st>>>(tp);
        g_launches.fetch_add(1);
        GSB_CUDA(cudaGetLastError());
    }
    SkParams sp;
    std::memset(&sp, 0, sizeof sp);
    sp.ctab = tp.ctab;
    sp.ttab = tp.ttab;
    sp.btile = tp.btile;
    sp.n_ytiles = L.n_ytiles;
    sp.lyp = L.lyp;
    sp.vrows = L.n_slow * L.lyp;
    sp.n_vt = L.n_vt;
    sp.n_col_tiles = L.n_col_tiles;
    sp.n_stages = n_stages;
    sp.ncomp = ncomp;
    sp.n_slow = L.n_slow;
    sp.n_in = L.n_in;
    sp.ly = L.ly;
    sp.lc = L.lc;
    sp.n_modes_pad = n_modes_pad;
    sp.out = d_out;
    sp.out_fstride = d_fstride;
    sp.epi = epi;
    const bool partial = (L.ly % SK_TM) != 0 || (L.lc % SK_TN) != 0;
    sp.slots = slots;
    unsigned *d_flags = tp.flags;
    std::vector<SkBound> bnd;
    for (int64_t k = 0; k < pieces; ++k) {
        const int64_t g0 = cuts[(size_t)k], g1 = cuts[(size_t)k + 1];
        if (g1 == g0) continue;
        // equal-cost shares of the tiles of this piece; the cost period is one slow index (nothing packed: every slow
        // index has the same row tiles) or one field (packed: all row tiles full except the last)
        const int grid = L.pack ? sk_plan(g0 * L.n_col_tiles, g1 * L.n_col_tiles, (int)L.n_vt, L.n_col_tiles, n_stages,
                                          L.n_slow * L.lyp, L.lc, max_grid, bnd)
                                : sk_plan(g0 * L.n_col_tiles, g1 * L.n_col_tiles, L.n_ytiles, L.n_col_tiles, n_stages, L.ly,
                                          L.lc, max_grid, bnd);
        for (int c = 0; c <= grid; ++c) sp.bnd[c] = bnd[(size_t)c];
        sp.flags = d_flags + k * max_grid;
        {
            KernelTimer timer(st);
            TraceScope ts("contract(sk)", st);
            GSB_TRY(sk_launch(sp, grid, scale, partial, L.pack, st));
        }
        const int64_t u0 = done_units(g0), u1 = (g1 == total_vt) ? n_z * L.n_slow : done_units(g1);
        if (h_out && u1 > u0) {
            cudaEvent_t ev = dev.contract_events[k % DeviceState::N_CHUNK_EVENTS];
            GSB_CUDA(cudaEventRecord(ev, st));
            GSB_CUDA(cudaStreamWaitEvent(dev.streams[1], ev, 0));
            // units [u0, u1): unit u = z * n_slow + slow lies at field z, offset slow * ly * lc (pieces of a mesh with
            // inner slow axes are whole fields).  One copy when both sides use the same field stride, else one per field
            const int64_t unit_elems = L.ly * L.lc;
            TraceScope tc("d2h piece", dev.streams[1]);
            if (d_fstride == h_fstride && d_fstride == mesh.n) {
                const size_t off = (size_t)u0 * unit_elems;
                GSB_CUDA(cudaMemcpyAsync(h_out + off, d_out + off, sizeof(double) * (size_t)(u1 - u0) * unit_elems,
                                         cudaMemcpyDeviceToHost, dev.streams[1]));
            } else {
                for (int64_t z = u0 / L.n_slow; z * L.n_slow < u1; ++z) {
                    const int64_t a = std::max(u0, z * L.n_slow) - z * L.n_slow;
                    const int64_t b = std::min(u1, (z + 1) * L.n_slow) - z * L.n_slow;
                    GSB_CUDA(cudaMemcpyAsync(h_out + z * h_fstride + a * unit_elems, d_out + z * d_fstride + a * unit_elems,
                                             sizeof(double) * (size_t)(b - a) * unit_elems, cudaMemcpyDeviceToHost,
                                             dev.streams[1]));
                }
            }
        }
    }
    if (h_out) {
        GSB_CUDA(cudaEventRecord(dev.events[4], dev.streams[1]));
        GSB_CUDA(cudaStreamWaitEvent(st, dev.events[4], 0));
    }
    g_cnt_sk.fetch_add(1);
    if (L.pack) g_cnt_packed.fetch_add(1);
    return GSB_OK;
}

// everything on device: d_cov (B,dim,N), d_z1/d_z2 (B,N), d_axes, d_out (B,ncomp,n).
// `h_out`: when non-null, finished pieces are copied to this host buffer as they complete.
// `st` is the caller's stream: all work is ordered after what `st` holds on entry, and `st` waits
// for all of it before this function returns.
static int structured_on_device(const double *d_cov, const double *d_z1, const double *d_z2,
                                const double *d_sf, const double *d_axes, const MeshInfo &mesh, int64_t n_modes,
                                int64_t n_batch, bool vec, const Epi &epi, double *d_out, int64_t d_fstride,
                                double *h_out, int64_t h_fstride, DeviceState &dev, cudaStream_t st)
{
    const int dim = mesh.dim;
    const int ncomp = vec ? dim : 1;
    Scratch scr(st);
    const int64_t force = g_opt_force_path.load();
    // Which path?  Estimated times from the same cost model that balances the stream-K shares:
    //   separable  contraction of the padded tiles + table building, with the last axis alone as column axis or
    //              -- thin meshes (1000 x 1000 x 10: reservoir layers) -- the last TWO axes folded into one
    //              column axis of len_y * len_z entries (its table holds exp(i (k'_y y + k'_z z)); C order
    //              makes the folded axis contiguous in the output);
    //   direct     expand the mesh on the device, D + 13 FP64 instructions per pair at ~86 % of the pipe.
    const int64_t n_modes_pad = (n_modes + SK_KC - 1) / SK_KC * SK_KC;
    bool separable = false;
    SkLayout layout;
    std::memset(&layout, 0, sizeof layout);
    if (dim >= 2 && n_modes > 0) {
        double t_best = 1e300;
        const int64_t fold_opt = g_opt_fold_axes.load();
        for (int fold = 0; fold < 2; ++fold) {
            if (fold == 1) {
                const int64_t wc2 = mesh.len[dim - 2] * mesh.len[dim - 1];
                const bool fits = dim >= 3 && wc2 * n_modes_pad * ncomp * n_batch <= ((int64_t)1 << 28);   // <= 4 GiB
                if (!fits || fold_opt == 0) continue;
            } else if (fold_opt == 2 && dim >= 3 &&
                       mesh.len[dim - 2] * mesh.len[dim - 1] * n_modes_pad * ncomp * n_batch <= ((int64_t)1 << 28)) {
                continue;     // forced folding
            }
            double t = 0.0;
            const SkLayout cand = sk_choose_layout(mesh, n_modes_pad, n_batch, ncomp, fold == 1, dev.sm_count,
                                                   h_out != nullptr, &t);
            if (t < t_best) { t_best = t; layout = cand; }
        }
        const double instr = (double)(dim + 13 + (vec ? dim : 0));
        const double t_direct = (double)mesh.n * (double)n_modes * (double)n_batch * instr /
                                    ((double)dev.sm_count * 64.0 * 1.9e9 * 0.86) + 12e-6 * (double)n_batch +
                                (h_out ? (double)mesh.n * (double)(n_batch * ncomp) * sizeof(double) / 50e9 : 0.0);
        separable = t_best < t_direct;
        const int64_t tiles = ((mesh.n_rows + SK_TM - 1) / SK_TM) * ((mesh.len[dim - 1] + SK_TN - 1) / SK_TN) * n_batch * ncomp;
        if (tiles < g_opt_structured_min_tiles.load()) separable = false;
        if (force == 1) separable = false;
        if (force == 2) separable = true;
    }

    if (!separable) {
        // expand on the device, then the direct kernel per batch entry
        ExpandParams ep;
        double *d_pos = nullptr;
        GSB_TRY(scr.alloc(&d_pos, (size_t)dim * mesh.n));
        ep.axes = d_axes;
        for (int t = 0; t < dim; ++t) {
            ep.axis_off[t] = mesh.off[t];
            ep.axis_len[t] = mesh.len[t];
        }
        std::memcpy(ep.matrix, mesh.matrix, sizeof ep.matrix);
        ep.dim = dim;
        ep.n = mesh.n;
        ep.pos = d_pos;
        const int blocks = (int)std::min<int64_t>((mesh.n + 255) / 256, 8 * (int64_t)dev.sm_count);
        expand_grid_kernel<<<std::max(blocks, 1), 256, 0, st>>>(ep);
        g_launches.fetch_add(1);
        GSB_CUDA(cudaGetLastError());
        for (int64_t b = 0; b < n_batch; ++b) {
            double *d_recs = nullptr;
            int64_t pad = 0;
            GSB_TRY(pack_modes(d_cov + b * dim * n_modes, d_z1 + b * n_modes, d_z2 + b * n_modes,
                               d_sf ? d_sf + b * n_modes : nullptr, dim, n_modes, vec, &d_recs, &pad, scr, st));
            GSB_TRY(direct_on_device(d_recs, pad, d_pos, mesh.n, dim, vec, mesh.n,
                                     d_out + b * ncomp * d_fstride, d_fstride, epi, dev, scr, st));
        }
        if (h_out) {
            GSB_CUDA(cudaMemcpy2DAsync(h_out, sizeof(double) * h_fstride, d_out, sizeof(double) * d_fstride,
                                       sizeof(double) * mesh.n, (size_t)(n_batch * ncomp), cudaMemcpyDeviceToHost, st));
        }
        return GSB_OK;
    }

    GSB_TRY(sk_on_device(d_cov, d_z1, d_z2, d_sf, d_axes, mesh, layout, n_modes, n_batch, vec, epi, d_out, d_fstride,
                         h_out, h_fstride, dev, st));
    g_cnt_separable.fetch_add(1);
    if (layout.n_col_axes == 2) g_cnt_folded.fetch_add(1);
    if (g_opt_trace.load() && !g_trace.empty()) {
        cudaDeviceSynchronize();
        for (auto &r : g_trace) {
            float a = 0.f, b = 0.f;
            cudaEventElapsedTime(&a, g_trace[0].e0, r.e0);
            cudaEventElapsedTime(&b, g_trace[0].e0, r.e1);
            fprintf(stderr, "[gsb trace] %-12s %8.3f -> %8.3f ms\n", r.name, a, b);
        }
        for (auto &r : g_trace) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
        g_trace.clear();
    }
    return GSB_OK;
}

static int structured_impl(const double *cov, const double *z1, const double *z2, const double *sf,
                           const double *axes, const int64_t *axis_len, const double *matrix, int dim,
                           int64_t n_modes, int64_t n_batch, double *out, bool vec,
                           const gsb_epilogue *epilogue, int mem, int device, void *stream,
                           const gsb_point_epilogue *pepi = nullptr, const SlabOpt *slab = nullptr)
{
    DeviceGuard guard;
    GSB_TRY(check_common(cov, z1, z2, dim, n_modes));
    GSB_TRY(check_epilogue(epilogue));
    GSB_TRY(check_point_epilogue(pepi, vec));
    Epi epi = make_epi(epilogue, pepi);
    if (!axis_len) return fail(GSB_ERR_ARGUMENT, "axis_len must not be NULL");
    if (n_batch < 1) return fail(GSB_ERR_ARGUMENT, "n_batch must be >= 1");
    if (vec && dim != 2 && dim != 3)
        return fail(GSB_ERR_ARGUMENT, "summate_incompr: dim must be 2 or 3 (generator.py:514-517)");
    if (mem != GSB_MEM_HOST && mem != GSB_MEM_DEVICE)
        return fail(GSB_ERR_ARGUMENT, "mem must be GSB_MEM_HOST or GSB_MEM_DEVICE");
    MeshInfo mesh;
    mesh.dim = dim;
    mesh.total_axes = 0;
    mesh.n = 1;
    for (int t = 0; t < dim; ++t) {
        if (axis_len[t] < 0) return fail(GSB_ERR_ARGUMENT, "axis_len must be >= 0");
        mesh.len[t] = axis_len[t];
        mesh.off[t] = mesh.total_axes;
        mesh.total_axes += axis_len[t];
        mesh.n *= axis_len[t];
    }
    if (mesh.n == 0) return GSB_OK;
    if (!axes || !out) return fail(GSB_ERR_ARGUMENT, "axes and out must not be NULL");
    // a slab evaluates the entries [lo, hi) of axis 0; `out` and the per-point arrays still describe the FULL mesh
    const int64_t fstride = mesh.n;
    if (slab) {
        if (slab->lo < 0 || slab->hi < slab->lo || slab->hi > mesh.len[0])
            return fail(GSB_ERR_ARGUMENT, "slab outside axis 0");
        const int64_t rest = mesh.n / mesh.len[0];
        mesh.off[0] += slab->lo;
        mesh.len[0] = slab->hi - slab->lo;
        mesh.n = mesh.len[0] * rest;
        if (mesh.n == 0) return GSB_OK;
        out += slab->lo * rest;
        epi = epi_shift(epi, slab->lo * rest);
    }
    mesh.n_rows = mesh.n / mesh.len[dim - 1];
    std::memset(mesh.matrix, 0, sizeof mesh.matrix);
    mesh.identity = (matrix == nullptr);
    // the matrix is dim*dim doubles and is always read on the host (it is tiny)
    std::vector<double> hmat((size_t)dim * dim, 0.0);
    if (matrix) {
        if (mem == GSB_MEM_DEVICE) {
            DeviceState *dv = nullptr;
            GSB_TRY(ensure_device(device, &dv));
            cudaPointerAttributes attr;
            cudaError_t pe = cudaPointerGetAttributes(&attr, matrix);
            if (pe == cudaSuccess && (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged)) {
                GSB_CUDA(cudaMemcpyAsync(hmat.data(), matrix, sizeof(double) * dim * dim,
                                         cudaMemcpyDeviceToHost, static_cast<cudaStream_t>(stream)));
                GSB_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
            } else {
                cudaGetLastError();
                std::memcpy(hmat.data(), matrix, sizeof(double) * dim * dim);
            }
        } else {
            std::memcpy(hmat.data(), matrix, sizeof(double) * dim * dim);
        }
    } else {
        for (int t = 0; t < dim; ++t) hmat[(size_t)t * dim + t] = 1.0;
    }
    for (int t = 0; t < dim * dim; ++t) mesh.matrix[t] = hmat[t];

    DeviceState *dev = nullptr;
    GSB_TRY(ensure_device(device, &dev));
    const int ncomp = vec ? dim : 1;

    if (mem == GSB_MEM_DEVICE) {
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        std::lock_guard<std::mutex> lock(dev->call_mutex);
        return structured_on_device(cov, z1, z2, sf, axes, mesh, n_modes, n_batch, vec, epi, out, fstride, nullptr, 0,
                                    *dev, st);
    }
    std::lock_guard<std::mutex> lock(dev->call_mutex);
    cudaStream_t s0 = dev->streams[0];
    Scratch scr(s0);
    double *d_cov, *d_z1, *d_z2, *d_axes, *d_out, *d_sf = nullptr;
    if (sf) {
        GSB_TRY(scr.alloc(&d_sf, (size_t)n_batch * n_modes));
        if (n_modes > 0)
            GSB_CUDA(cudaMemcpyAsync(d_sf, sf, sizeof(double) * n_batch * n_modes, cudaMemcpyHostToDevice, s0));
    }
    GSB_TRY(scr.alloc(&d_cov, (size_t)n_batch * dim * n_modes));
    GSB_TRY(scr.alloc(&d_z1, (size_t)n_batch * n_modes));
    GSB_TRY(scr.alloc(&d_z2, (size_t)n_batch * n_modes));
    GSB_TRY(scr.alloc(&d_axes, (size_t)mesh.total_axes));
    GSB_TRY(scr.alloc(&d_out, (size_t)n_batch * ncomp * mesh.n));
    if (n_modes > 0) {
        GSB_CUDA(cudaMemcpyAsync(d_cov, cov, sizeof(double) * n_batch * dim * n_modes, cudaMemcpyHostToDevice, s0));
        GSB_CUDA(cudaMemcpyAsync(d_z1, z1, sizeof(double) * n_batch * n_modes, cudaMemcpyHostToDevice, s0));
        GSB_CUDA(cudaMemcpyAsync(d_z2, z2, sizeof(double) * n_batch * n_modes, cudaMemcpyHostToDevice, s0));
    }
    GSB_CUDA(cudaMemcpyAsync(d_axes, axes, sizeof(double) * mesh.total_axes, cudaMemcpyHostToDevice, s0));
    GSB_TRY(structured_on_device(d_cov, d_z1, d_z2, d_sf, d_axes, mesh, n_modes, n_batch, vec, epi, d_out, mesh.n, out,
                                 fstride, *dev, s0));
    GSB_CUDA(cudaStreamSynchronize(s0));
    return GSB_OK;
}

// ---------------------------------------------------------------------------------------------
// kriging evaluation (row f1)
// ---------------------------------------------------------------------------------------------
static std::atomic<int64_t> g_cnt_krige{0};
static std::atomic<int64_t> g_opt_krige_host_chunk_mb{256};

struct KrigeOperand {
    double *w = nullptr;       // M^T cond
    double *atile = nullptr;   // pre-tiled [L; w]
    double *zeros = nullptr;
    int K = 0, R = 0;
};

// d_mat (K,K), d_cond (K,) on the device
static int krige_prepare(const double *d_mat, const double *d_cond, int K, bool want_var, KrigeOperand *op,
                         Scratch &scr, cudaStream_t st)
{
    op->K = K;
    op->R = (K + 1 + SEP_TM - 1) / SEP_TM;
    GSB_TRY(scr.alloc(&op->w, (size_t)K));
    krige_w_kernel<<<(K + 127) / 128, 128, 0, st>>>(d_mat, d_cond, K, op->w);
    g_launches.fetch_add(1);
    GSB_CUDA(cudaGetLastError());
    if (!want_var) return GSB_OK;
    const int64_t n_tiles = krige_tile_off(op->R);
    GSB_TRY(scr.alloc(&op->atile, (size_t)n_tiles * SEP_A_TILE));
    GSB_TRY(scr.alloc(&op->zeros, (size_t)SEP_TN));
    GSB_CUDA(cudaMemsetAsync(op->zeros, 0, sizeof(double) * SEP_TN, st));
    krige_tiles_kernel<<<(unsigned)n_tiles, 256, 0, st>>>(d_mat, op->w, K, op->R, op->atile);
    g_launches.fetch_add(1);
    GSB_CUDA(cudaGetLastError());
    return GSB_OK;
}

// d_kv (K, n) with row stride ld on the device -> d_field, d_error (n,)
// `pad_ok`: for odd n the caller guarantees that column n of every row is readable and finite (ld > n).
static int krige_on_device(const KrigeOperand &op, const double *d_kv, int64_t ld, int64_t n, bool pad_ok,
                           double *d_field, double *d_error, const DeviceState &dev, Scratch &scr, cudaStream_t st)
{
    if (n == 0) return GSB_OK;
    const int K = op.K;
    if (!d_error) {
        krige_field_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(op.w, d_kv, ld, K, n, d_field);
        g_launches.fetch_add(1);
        GSB_CUDA(cudaGetLastError());
        g_cnt_krige.fetch_add(1);
        return GSB_OK;
    }
    KrigeParams kp;
    std::memset(&kp, 0, sizeof kp);
    kp.atile = op.atile;
    kp.K = K;
    kp.R = op.R;
    kp.n_pairs = (op.R + 1) / 2;
    kp.n_dstages = (K + KRG_KD - 1) / KRG_KD;
    kp.zeros = op.zeros;
    const bool aligned = (reinterpret_cast<uintptr_t>(d_kv) & 15) == 0 && (ld & 1) == 0 &&
                         ((n & 1) == 0 || (pad_ok && ld > n));
    // column chunks: bound the partial sums (and the repack buffer when the caller's array cannot be
    // read by 16-byte bulk copies)
    int64_t chunk = n;
    if (!aligned) {
        chunk = std::max<int64_t>(SEP_TN, ((int64_t)256 << 20) / ((int64_t)K * 8) / SEP_TN * SEP_TN);
        chunk = std::min<int64_t>(chunk, (n + 1) / 2 * 2);
    }
    double *d_pack = nullptr, *d_partial = nullptr;
    if (!aligned) GSB_TRY(scr.alloc(&d_pack, (size_t)K * chunk));
    GSB_TRY(scr.alloc(&d_partial, (size_t)kp.n_pairs * std::min<int64_t>(chunk, n)));
    for (int64_t c0 = 0; c0 < n; c0 += chunk) {
        const int64_t m = std::min(chunk, n - c0);
        if (aligned) {
            kp.kv = d_kv + c0;
            kp.ld = ld;
            kp.n = m;
            kp.n_copy = (m + 1) / 2 * 2;
        } else {
            const int64_t m_pad = (m + 1) / 2 * 2;
            const int blocks = (int)std::min<int64_t>(((int64_t)K * m_pad + 255) / 256, 32LL * dev.sm_count);
            krige_repack_kernel<<<blocks, 256, 0, st>>>(d_kv + c0, ld, K, m, d_pack, m_pad);
            g_launches.fetch_add(1);
            GSB_CUDA(cudaGetLastError());
            kp.kv = d_pack;
            kp.ld = m_pad;
            kp.n = m;
            kp.n_copy = m_pad;   // the bulk copies read m rounded up to even: the zero padding column
        }
        kp.n_col_tiles = (m + SEP_TN - 1) / SEP_TN;
        kp.partial = d_partial;
        kp.field = d_field + c0;
        {
            KernelTimer timer(st);
            GSB_TRY(launch_krige(kp, false, dev.sm_count, st));
        }
        const int blocks = (int)std::min<int64_t>((m + 255) / 256, 8LL * dev.sm_count);
        krige_finish_kernel<<<blocks, 256, 0, st>>>(d_partial, kp.n_pairs, m, d_error + c0);
        g_launches.fetch_add(1);
        GSB_CUDA(cudaGetLastError());
    }
    g_cnt_krige.fetch_add(1);
    return GSB_OK;
}

static int krige_impl(const double *mat, const double *kv, int64_t ld, const double *cond, int64_t K64, int64_t n,
                      double *field, double *error, bool want_var, int mem, int device, void *stream)
{
    DeviceGuard guard;
    if (K64 < 0 || K64 > (1 << 20)) return fail(GSB_ERR_ARGUMENT, "krige: K out of range");
    if (n < 0) return fail(GSB_ERR_ARGUMENT, "krige: n must be >= 0");
    if (mem != GSB_MEM_HOST && mem != GSB_MEM_DEVICE)
        return fail(GSB_ERR_ARGUMENT, "mem must be GSB_MEM_HOST or GSB_MEM_DEVICE");
    if (n == 0) return GSB_OK;
    if (!field || (want_var && !error)) return fail(GSB_ERR_ARGUMENT, "krige: output pointers must not be NULL");
    if (K64 > 0 && (!mat || !kv || !cond)) return fail(GSB_ERR_ARGUMENT, "krige: krig_mat, krig_vecs and cond must not be NULL");
    if (ld < n) return fail(GSB_ERR_ARGUMENT, "krige: leading dimension smaller than n");
    const int K = (int)K64;
    DeviceState *dev = nullptr;
    GSB_TRY(ensure_device(device, &dev));
    if (K == 0) {   // empty system: both sums are empty
        if (mem == GSB_MEM_DEVICE) {
            cudaStream_t st = static_cast<cudaStream_t>(stream);
            GSB_CUDA(cudaMemsetAsync(field, 0, sizeof(double) * n, st));
            if (want_var) GSB_CUDA(cudaMemsetAsync(error, 0, sizeof(double) * n, st));
        } else {
            std::memset(field, 0, sizeof(double) * n);
            if (want_var) std::memset(error, 0, sizeof(double) * n);
        }
        return GSB_OK;
    }
    if (mem == GSB_MEM_DEVICE) {
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        Scratch scr(st);
        KrigeOperand op;
        GSB_TRY(krige_prepare(mat, cond, K, want_var, &op, scr, st));
        return krige_on_device(op, kv, ld, n, false, field, want_var ? error : nullptr, *dev, scr, st);
    }
    // ---- host buffers: operand once, then column chunks of krig_vecs double buffered over two streams ----
    std::lock_guard<std::mutex> lock(dev->call_mutex);
    cudaStream_t s0 = dev->streams[0];
    Scratch scr0(s0), scr1(dev->streams[1]);
    double *d_mat, *d_cond;
    GSB_TRY(scr0.alloc(&d_mat, (size_t)K * K));
    GSB_TRY(scr0.alloc(&d_cond, (size_t)K));
    GSB_CUDA(cudaMemcpyAsync(d_mat, mat, sizeof(double) * K * K, cudaMemcpyHostToDevice, s0));
    GSB_CUDA(cudaMemcpyAsync(d_cond, cond, sizeof(double) * K, cudaMemcpyHostToDevice, s0));
    KrigeOperand op;
    GSB_TRY(krige_prepare(d_mat, d_cond, K, want_var, &op, scr0, s0));
    GSB_CUDA(cudaEventRecord(dev->events[0], s0));
    GSB_CUDA(cudaStreamWaitEvent(dev->streams[1], dev->events[0], 0));
    int64_t chunk = std::max<int64_t>(SEP_TN, (g_opt_krige_host_chunk_mb.load() << 20) / ((int64_t)K * 8) / SEP_TN * SEP_TN);
    chunk = std::min<int64_t>(chunk, (n + 1) / 2 * 2);
    const int nbuf = n > chunk ? 2 : 1;
    double *d_kv[2] = {nullptr, nullptr}, *d_f[2] = {nullptr, nullptr}, *d_e[2] = {nullptr, nullptr};
    for (int b = 0; b < nbuf; ++b) {
        Scratch &s = b ? scr1 : scr0;
        GSB_TRY(s.alloc(&d_kv[b], (size_t)K * chunk));
        GSB_TRY(s.alloc(&d_f[b], (size_t)chunk));
        if (want_var) GSB_TRY(s.alloc(&d_e[b], (size_t)chunk));
    }
    int64_t c = 0;
    for (int64_t c0 = 0; c0 < n; c0 += chunk, ++c) {
        const int b = (int)(c % nbuf);
        cudaStream_t st = dev->streams[b];
        Scratch &s = b ? scr1 : scr0;
        const int64_t m = std::min(chunk, n - c0);
        if (m & 1)   // the bulk copies read an even number of columns: keep the padding column finite
            GSB_CUDA(cudaMemset2DAsync(d_kv[b] + m, sizeof(double) * chunk, 0, sizeof(double), K, st));
        GSB_CUDA(cudaMemcpy2DAsync(d_kv[b], sizeof(double) * chunk, kv + c0, sizeof(double) * ld, sizeof(double) * m, K,
                                   cudaMemcpyHostToDevice, st));
        // chunk buffers are aligned with an even row stride; an odd tail is padded with the zero column
        GSB_TRY(krige_on_device(op, d_kv[b], chunk, m, true, d_f[b], want_var ? d_e[b] : nullptr, *dev, s, st));
        GSB_CUDA(cudaMemcpyAsync(field + c0, d_f[b], sizeof(double) * m, cudaMemcpyDeviceToHost, st));
        if (want_var) GSB_CUDA(cudaMemcpyAsync(error + c0, d_e[b], sizeof(double) * m, cudaMemcpyDeviceToHost, st));
    }
    GSB_CUDA(cudaStreamSynchronize(dev->streams[1]));
    GSB_CUDA(cudaStreamSynchronize(s0));
    return GSB_OK;
}

// ---- whole evaluation loop on the device: right-hand sides generated, then contracted ----
struct KrigeEvalArgs {
    const gsb_cov_model *model;
    const double *mat, *cond, *cond_pos;
    int64_t K, C;
    int dim;
    const double *pos;       // flat variant
    int64_t pos_ld, n;
    const double *axes;      // structured variant
    const int64_t *axis_len;
    const double *matrix;
    int unbiased;
    const double *tail;
    int64_t tail_ld;
    double *field, *error;
    // share of a multi-GPU plan: evaluate the points [first, first + count) only; pos / tail / field / error still
    // describe (and are indexed by) the FULL point set.  count < 0: all points.
    int64_t first = 0, count = -1;
};

template <int D>
static void launch_kvgen(const KvgenParams &gp, int n_dstages, int64_t n_ct, cudaStream_t st)
{
    dim3 grid((unsigned)n_ct, (unsigned)std::min(8, std::max(1, n_dstages / 4)));
    kvgen_kernel<D><<<grid, 256, 0, st>>>(gp);
}

template <int D>
static int launch_field_gen(const KvgenParams &gp, cudaStream_t st)
{
    krige_field_gen_kernel<D><<<(unsigned)((gp.n + 255) / 256), 256, 0, st>>>(gp);
    return GSB_OK;
}

// all pointers on the device except axis_len / matrix (host, tiny)
static int krige_eval_on_device(const KrigeEvalArgs &a, const MeshInfo *mesh, const DeviceState &dev, cudaStream_t st)
{
    const int K = (int)a.K, C = (int)a.C, D = a.dim;
    Scratch scr(st);
    KrigeOperand op;
    GSB_TRY(krige_prepare(a.mat, a.cond, K, a.error != nullptr, &op, scr, st));
    KvgenParams gp;
    std::memset(&gp, 0, sizeof gp);
    gp.cov.type = a.model->type;
    gp.cov.exact = a.model->exact;
    gp.cov.var = a.model->var;
    gp.cov.len_rescaled = a.model->len_rescaled;
    gp.cov.sill = a.model->sill;
    gp.cov.param = a.model->param;
    gp.dim = D;
    gp.C = C;
    gp.K = K;
    gp.unbiased = a.unbiased ? 1 : 0;
    gp.n_dstages = (K + KRG_KD - 1) / KRG_KD;
    gp.cond_pos = a.cond_pos;
    gp.pos = a.pos;
    gp.pos_ld = a.pos_ld;
    if (mesh) {
        gp.axes = a.axes;
        for (int t = 0; t < D; ++t) {
            gp.axis_off[t] = mesh->off[t];
            gp.axis_len[t] = mesh->len[t];
        }
        std::memcpy(gp.matrix, mesh->matrix, sizeof gp.matrix);
    }
    gp.tail = a.tail;
    gp.tail_ld = a.tail_ld;
    gp.w = op.w;
    const int64_t first = a.count < 0 ? 0 : a.first;
    const int64_t n = a.count < 0 ? a.n : a.count;
#define GSB_KRG_DIM_SWITCH(CALL)                                                              \
    switch (D) {                                                                              \
    case 1: CALL(1); break;                                                                   \
    case 2: CALL(2); break;                                                                   \
    case 3: CALL(3); break;                                                                   \
    case 4: CALL(4); break;                                                                   \
    default: return fail(GSB_ERR_ARGUMENT, "krige_evaluate: dim must be in 1..4");            \
    }
    if (!a.error) {
        gp.col_begin = first;
        gp.n = n;
        gp.field = a.field + first;
#define GSB_CALL(DD) GSB_TRY(launch_field_gen<DD>(gp, st))
        GSB_KRG_DIM_SWITCH(GSB_CALL)
#undef GSB_CALL
        g_launches.fetch_add(1);
        GSB_CUDA(cudaGetLastError());
        g_cnt_krige.fetch_add(1);
        return GSB_OK;
    }
    KrigeParams kp;
    std::memset(&kp, 0, sizeof kp);
    kp.atile = op.atile;
    kp.K = K;
    kp.R = op.R;
    kp.n_pairs = (op.R + 1) / 2;
    kp.n_dstages = gp.n_dstages;
    kp.zeros = op.zeros;
    // column chunks sized by the scratch budget (half of "scratch_mb" for the tiled right-hand sides)
    const size_t col_tile_bytes = (size_t)gp.n_dstages * SEP_B_TILE * sizeof(double);
    int64_t chunk_tiles = std::max<int64_t>(1, (int64_t)((size_t)g_opt_scratch_mb.load() * (1u << 20) / 2 / col_tile_bytes));
    chunk_tiles = std::min<int64_t>(chunk_tiles, (n + SEP_TN - 1) / SEP_TN);
    double *d_btile = nullptr, *d_partial = nullptr;
    GSB_TRY(scr.alloc(&d_btile, (size_t)chunk_tiles * gp.n_dstages * SEP_B_TILE));
    GSB_TRY(scr.alloc(&d_partial, (size_t)kp.n_pairs * chunk_tiles * SEP_TN));
    gp.btile = d_btile;
    kp.btile = d_btile;
    kp.partial = d_partial;
    for (int64_t c0 = 0; c0 < n; c0 += chunk_tiles * SEP_TN) {
        const int64_t m = std::min<int64_t>(chunk_tiles * SEP_TN, n - c0);
        const int64_t n_ct = (m + SEP_TN - 1) / SEP_TN;
        gp.col_begin = first + c0;
        gp.n = m;
#define GSB_CALL(DD) launch_kvgen<DD>(gp, gp.n_dstages, n_ct, st)
        GSB_KRG_DIM_SWITCH(GSB_CALL)
#undef GSB_CALL
        g_launches.fetch_add(1);
        GSB_CUDA(cudaGetLastError());
        kp.n = m;
        kp.n_copy = m;
        kp.n_col_tiles = n_ct;
        kp.field = a.field + first + c0;
        {
            KernelTimer timer(st);
            GSB_TRY(launch_krige(kp, true, dev.sm_count, st));
        }
        const int blocks = (int)std::min<int64_t>((m + 255) / 256, 8LL * dev.sm_count);
        krige_finish_kernel<<<blocks, 256, 0, st>>>(d_partial, kp.n_pairs, m, a.error + first + c0);
        g_launches.fetch_add(1);
        GSB_CUDA(cudaGetLastError());
    }
#undef GSB_KRG_DIM_SWITCH
    g_cnt_krige.fetch_add(1);
    return GSB_OK;
}

static int krige_eval_impl(KrigeEvalArgs a, bool structured, int mem, int device, void *stream)
{
    DeviceGuard guard;
    if (!a.model) return fail(GSB_ERR_ARGUMENT, "krige_evaluate: model must not be NULL");
    if (a.model->type < GSB_COV_GAUSSIAN || a.model->type > GSB_COV_SPHERICAL)
        return fail(GSB_ERR_ARGUMENT, "krige_evaluate: unknown covariance model type");
    if (!(a.model->len_rescaled > 0.0)) return fail(GSB_ERR_ARGUMENT, "krige_evaluate: len_rescaled must be > 0");
    if (a.dim < 1 || a.dim > 4) return fail(GSB_ERR_ARGUMENT, "krige_evaluate: dim must be in 1..4");
    if (a.K < 1 || a.K > (1 << 20)) return fail(GSB_ERR_ARGUMENT, "krige_evaluate: krige_size out of range");
    a.unbiased = a.unbiased ? 1 : 0;
    if (a.C < 0 || a.C + a.unbiased > a.K) return fail(GSB_ERR_ARGUMENT, "krige_evaluate: cond_no + unbiased exceeds krige_size");
    const int64_t n_tail = a.K - a.C - a.unbiased;
    if (!a.mat || !a.cond || (a.C > 0 && !a.cond_pos)) return fail(GSB_ERR_ARGUMENT, "krige_evaluate: NULL input");
    if (mem != GSB_MEM_HOST && mem != GSB_MEM_DEVICE)
        return fail(GSB_ERR_ARGUMENT, "mem must be GSB_MEM_HOST or GSB_MEM_DEVICE");
    MeshInfo mesh;
    std::memset(&mesh, 0, sizeof mesh);
    if (structured) {
        if (!a.axis_len) return fail(GSB_ERR_ARGUMENT, "axis_len must not be NULL");
        mesh.dim = a.dim;
        mesh.n = 1;
        for (int t = 0; t < a.dim; ++t) {
            if (a.axis_len[t] < 0) return fail(GSB_ERR_ARGUMENT, "axis_len must be >= 0");
            mesh.len[t] = a.axis_len[t];
            mesh.off[t] = mesh.total_axes;
            mesh.total_axes += a.axis_len[t];
            mesh.n *= a.axis_len[t];
        }
        a.n = mesh.n;
        for (int t = 0; t < a.dim; ++t)
            for (int u = 0; u < a.dim; ++u)
                mesh.matrix[t * a.dim + u] = a.matrix ? a.matrix[t * a.dim + u] : (t == u ? 1.0 : 0.0);
    }
    if (a.n < 0) return fail(GSB_ERR_ARGUMENT, "krige_evaluate: n_pts must be >= 0");
    if (a.n == 0) return GSB_OK;
    if (a.count >= 0 && (a.first < 0 || a.first + a.count > a.n)) return fail(GSB_ERR_ARGUMENT, "krige_evaluate: share outside the point set");
    if (a.count == 0) return GSB_OK;
    if (!a.field) return fail(GSB_ERR_ARGUMENT, "krige_evaluate: field must not be NULL");
    if (structured ? !a.axes : (!a.pos || a.pos_ld < a.n)) return fail(GSB_ERR_ARGUMENT, "krige_evaluate: bad positions");
    if (n_tail > 0 && (!a.tail || a.tail_ld < a.n)) return fail(GSB_ERR_ARGUMENT, "krige_evaluate: drift rows missing");
    DeviceState *dev = nullptr;
    GSB_TRY(ensure_device(device, &dev));
    if (mem == GSB_MEM_DEVICE)
        return krige_eval_on_device(a, structured ? &mesh : nullptr, *dev, static_cast<cudaStream_t>(stream));

    std::lock_guard<std::mutex> lock(dev->call_mutex);
    cudaStream_t s0 = dev->streams[0];
    Scratch scr(s0);
    KrigeEvalArgs d = a;
    // this call's points: [first, first + cnt).  The device code indexes positions, drift rows and outputs by the
    // GLOBAL point index, so the staged copies of the share are addressed through pointers shifted by -first.
    const int64_t first = a.count < 0 ? 0 : a.first;
    const int64_t cnt = a.count < 0 ? a.n : a.count;
    d.first = first;
    d.count = cnt;
    double *p = nullptr;
    auto up = [&](const double *src, size_t count, const double **dst) -> int {
        GSB_TRY(scr.alloc(&p, count));
        if (count) GSB_CUDA(cudaMemcpyAsync(p, src, sizeof(double) * count, cudaMemcpyHostToDevice, s0));
        *dst = p;
        return GSB_OK;
    };
    GSB_TRY(up(a.mat, (size_t)a.K * a.K, &d.mat));
    GSB_TRY(up(a.cond, (size_t)a.K, &d.cond));
    GSB_TRY(up(a.cond_pos, (size_t)a.dim * a.C, &d.cond_pos));
    if (structured) {
        GSB_TRY(up(a.axes, (size_t)mesh.total_axes, &d.axes));
    } else {
        double *dp = nullptr;
        GSB_TRY(scr.alloc(&dp, (size_t)a.dim * cnt));
        GSB_CUDA(cudaMemcpy2DAsync(dp, sizeof(double) * cnt, a.pos + first, sizeof(double) * a.pos_ld, sizeof(double) * cnt,
                                   a.dim, cudaMemcpyHostToDevice, s0));
        d.pos = dp - first;
        d.pos_ld = cnt;
    }
    if (n_tail > 0) {
        double *dt = nullptr;
        GSB_TRY(scr.alloc(&dt, (size_t)n_tail * cnt));
        GSB_CUDA(cudaMemcpy2DAsync(dt, sizeof(double) * cnt, a.tail + first, sizeof(double) * a.tail_ld,
                                   sizeof(double) * cnt, n_tail, cudaMemcpyHostToDevice, s0));
        d.tail = dt - first;
        d.tail_ld = cnt;
    }
    double *df = nullptr, *de = nullptr;
    GSB_TRY(scr.alloc(&df, (size_t)cnt));
    d.field = df - first;
    if (a.error) {
        GSB_TRY(scr.alloc(&de, (size_t)cnt));
        d.error = de - first;
    }
    GSB_TRY(krige_eval_on_device(d, structured ? &mesh : nullptr, *dev, s0));
    GSB_CUDA(cudaMemcpyAsync(a.field + first, df, sizeof(double) * cnt, cudaMemcpyDeviceToHost, s0));
    if (a.error) GSB_CUDA(cudaMemcpyAsync(a.error + first, de, sizeof(double) * cnt, cudaMemcpyDeviceToHost, s0));
    GSB_CUDA(cudaStreamSynchronize(s0));
    return GSB_OK;
}

// ---------------------------------------------------------------------------------------------
// epilogue + microbenchmarks
// ---------------------------------------------------------------------------------------------
__global__ void scale_shift_kernel(double *f, int64_t n, double scale, double shift)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        f[i] = fma(scale, f[i], shift);
}

// CondSRF.get_scaling without nugget (cond_srf.py:175-177) after the clamp of krige/base.py:296-298.
// numpy's maximum(x, 0) keeps x when x >= 0 or x is NaN; IEEE sub / div / sqrt give numpy's bits.
__global__ void cond_scaling_kernel(const double *__restrict__ error, int64_t n, double sill, double var,
                                    double *krige_var, double *__restrict__ gain)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const double d = __dsub_rn(sill, error[i]);
        const double kv = (d >= 0.0 || d != d) ? d : 0.0;
        if (krige_var) krige_var[i] = kv;
        gain[i] = __dsqrt_rn(__ddiv_rn(kv, var));
    }
}

// 8 independent DFMA chains per thread, register resident
__global__ void __launch_bounds__(256) dfma_peak_kernel(double *sink, int iters, double seed)
{
    double a0 = seed, a1 = seed + 1, a2 = seed + 2, a3 = seed + 3;
    double a4 = seed + 4, a5 = seed + 5, a6 = seed + 6, a7 = seed + 7;
    const double m = 0.999999, c = 1e-7 * threadIdx.x;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
            a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
        }
    }
    const double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 12345.678) sink[0] = s;  // never true; keeps the chains alive
}

// DMMA m8n8k4: 8 independent accumulator tiles per warp
__global__ void __launch_bounds__(256) dmma_peak_kernel(double *sink, int iters, double seed)
{
    double c[8][2];
#pragma unroll
    for (int t = 0; t < 8; ++t) c[t][0] = c[t][1] = seed + t;
    const double a = 0.999999 + 1e-9 * (threadIdx.x & 31), b = 1e-3;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                asm volatile(
                    "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                    : "+d"(c[t][0]), "+d"(c[t][1])
                    : "d"(a), "d"(b));
            }
        }
    }
    double s = 0.0;
#pragma unroll
    for (int t = 0; t < 8; ++t) s += c[t][0] + c[t][1];
    if (s == 12345.678) sink[0] = s;
}

// ---------------------------------------------------------------------------------------------
// multi-GPU plan: one host thread per device; a call is cut into independent
// shares (point ranges, axis-0 slabs, batch entries) -- no inter-GPU traffic during the sum
// ---------------------------------------------------------------------------------------------
struct PlanWorker {
    int device = 0;
    std::thread th;
    std::mutex m;
    std::condition_variable cv;
    std::function<int()> job;
    bool pending = false, finished = true, quit = false;
    int rc = GSB_OK;
    std::string err;
    cudaEvent_t begin_ev = nullptr;   // recorded on the caller's stream when this device is the home device
    cudaEvent_t done_ev = nullptr;    // device route: this device's share is enqueued up to here

    void loop()
    {
        cudaSetDevice(device);
        std::unique_lock<std::mutex> lk(m);
        for (;;) {
            cv.wait(lk, [&] { return pending || quit; });
            if (quit) return;
            std::function<int()> j = std::move(job);
            pending = false;
            lk.unlock();
            const int r = j();
            std::string e = (r != GSB_OK) ? last_error_ref() : std::string();
            lk.lock();
            rc = r;
            err = std::move(e);
            finished = true;
            cv.notify_all();
        }
    }
    void submit(std::function<int()> j)
    {
        std::lock_guard<std::mutex> lk(m);
        job = std::move(j);
        pending = true;
        finished = false;
        cv.notify_all();
    }
    int wait()
    {
        std::unique_lock<std::mutex> lk(m);
        cv.wait(lk, [&] { return finished; });
        return rc;
    }
};

inline void plan_share(int64_t n, int parts, int part, int64_t *lo, int64_t *hi)
{
    const int64_t base = n / parts, rem = n % parts;
    *lo = part * base + std::min<int64_t>(part, rem);
    *hi = *lo + base + (part < rem ? 1 : 0);
}

}  // namespace gsb

struct gsb_plan {
    std::vector<std::unique_ptr<gsb::PlanWorker>> workers;
    std::mutex call_mutex;
    bool peer_ok = true;
};

namespace gsb {

// hand every device its job, wait for all, report the first failure on the calling thread
static int plan_run(gsb_plan *plan, const std::vector<std::function<int()>> &jobs)
{
    const size_t n = plan->workers.size();
    for (size_t g = 0; g < n; ++g)
        if (jobs[g]) plan->workers[g]->submit(jobs[g]);
    int rc = GSB_OK;
    std::string err;
    for (size_t g = 0; g < n; ++g) {
        if (!jobs[g]) continue;
        const int r = plan->workers[g]->wait();
        if (r != GSB_OK && rc == GSB_OK) {
            rc = r;
            err = "device " + std::to_string(plan->workers[g]->device) + ": " + plan->workers[g]->err;
        }
    }
    return rc == GSB_OK ? GSB_OK : fail(rc, err);
}

static int plan_home(const gsb_plan *plan, int mem, int home_device, int *home_idx)
{
    *home_idx = -1;
    if (mem != GSB_MEM_HOST && mem != GSB_MEM_DEVICE)
        return fail(GSB_ERR_ARGUMENT, "mem must be GSB_MEM_HOST or GSB_MEM_DEVICE");
    if (mem == GSB_MEM_HOST) return GSB_OK;
    for (size_t g = 0; g < plan->workers.size(); ++g)
        if (plan->workers[g]->device == home_device) *home_idx = (int)g;
    if (*home_idx < 0) return fail(GSB_ERR_ARGUMENT, "plan: home_device is not a device of the plan");
    if (!plan->peer_ok && plan->workers.size() > 1)
        return fail(GSB_ERR_ARGUMENT, "plan: no peer access between the plan's devices; use GSB_MEM_HOST");
    return GSB_OK;
}

// Device route of a plan call.  `share(g, st, scr)` enqueues device g's share on its stream `st` (scr: scratch that
// lives until the share has run).  The shares are ordered after everything `stream` (home device) holds on entry, and
// `stream` waits for all of them.
static int plan_run_device(gsb_plan *plan, int home_idx, void *stream,
                           const std::function<int(int, cudaStream_t, Scratch &)> &share)
{
    DeviceGuard guard;
    PlanWorker &home = *plan->workers[(size_t)home_idx];
    cudaStream_t user = static_cast<cudaStream_t>(stream);
    GSB_CUDA(cudaSetDevice(home.device));
    GSB_CUDA(cudaEventRecord(home.begin_ev, user));
    std::vector<std::function<int()>> jobs(plan->workers.size());
    for (size_t g = 0; g < plan->workers.size(); ++g) {
        PlanWorker *w = plan->workers[g].get();
        cudaEvent_t begin = home.begin_ev;
        jobs[g] = [w, g, begin, &share]() -> int {
            DeviceState *dev = nullptr;
            GSB_TRY(ensure_device(w->device, &dev));
            cudaStream_t st = dev->streams[2];
            GSB_CUDA(cudaStreamWaitEvent(st, begin, 0));
            {
                Scratch scr(st);
                GSB_TRY(share((int)g, st, scr));
            }
            GSB_CUDA(cudaEventRecord(w->done_ev, st));
            return GSB_OK;
        };
    }
    GSB_TRY(plan_run(plan, jobs));
    for (auto &w : plan->workers) GSB_CUDA(cudaStreamWaitEvent(user, w->done_ev, 0));
    return GSB_OK;
}

// a private copy of a small input array of the home device on the current device (peer reads of the mode set from
// every table entry would cross NVLink once per entry)
static int plan_stage(const double *src, size_t count, bool local, const double **dst, Scratch &scr, cudaStream_t st)
{
    if (local || !src || count == 0) {
        *dst = src;
        return GSB_OK;
    }
    double *p = nullptr;
    GSB_TRY(scr.alloc(&p, count));
    GSB_CUDA(cudaMemcpyAsync(p, src, sizeof(double) * count, cudaMemcpyDefault, st));
    *dst = p;
    return GSB_OK;
}

static int plan_summate_impl(gsb_plan *plan, const double *cov, const double *z1, const double *z2, const double *pos,
                             int64_t pos_ld, int dim, int64_t n_modes, int64_t n_pts, double *out, int64_t out_ld,
                             bool vec, const gsb_epilogue *epi, const gsb_point_epilogue *pepi, int mem,
                             int home_device, void *stream)
{
    if (!plan) return fail(GSB_ERR_ARGUMENT, "plan must not be NULL");
    std::lock_guard<std::mutex> lock(plan->call_mutex);
    const int G = (int)plan->workers.size();
    int home_idx = -1;
    GSB_TRY(plan_home(plan, mem, home_device, &home_idx));
    if (n_pts < 0) return fail(GSB_ERR_ARGUMENT, "n_pts must be >= 0");
    if (!vec) out_ld = n_pts;
    auto share_call = [=](int g, const double *c, const double *a, const double *b, int device, void *st) -> int {
        int64_t lo, hi;
        plan_share(n_pts, G, g, &lo, &hi);
        if (hi == lo) return GSB_OK;
        return summate_impl(c, a, b, nullptr, pos + lo, pos_ld, dim, n_modes, hi - lo, out + lo, out_ld, vec, epi, mem,
                            device, st, pepi ? pepi + g : nullptr, lo);
    };
    if (mem == GSB_MEM_HOST) {
        std::vector<std::function<int()>> jobs((size_t)G);
        for (int g = 0; g < G; ++g) {
            const int device = plan->workers[(size_t)g]->device;
            jobs[(size_t)g] = [=]() { return share_call(g, cov, z1, z2, device, nullptr); };
        }
        return plan_run(plan, jobs);
    }
    return plan_run_device(plan, home_idx, stream, [&](int g, cudaStream_t st, Scratch &scr) -> int {
        const bool local = (g == home_idx);
        const double *c, *a, *b;
        GSB_TRY(plan_stage(cov, (size_t)dim * n_modes, local, &c, scr, st));
        GSB_TRY(plan_stage(z1, (size_t)n_modes, local, &a, scr, st));
        GSB_TRY(plan_stage(z2, (size_t)n_modes, local, &b, scr, st));
        return share_call(g, c, a, b, plan->workers[(size_t)g]->device, st);
    });
}

static int plan_structured_impl(gsb_plan *plan, const double *cov, const double *z1, const double *z2,
                                const double *axes, const int64_t *axis_len, const double *matrix, int dim,
                                int64_t n_modes, int64_t n_batch, double *out, bool vec, const gsb_epilogue *epi,
                                const gsb_point_epilogue *pepi, int mem, int home_device, void *stream)
{
    if (!plan) return fail(GSB_ERR_ARGUMENT, "plan must not be NULL");
    std::lock_guard<std::mutex> lock(plan->call_mutex);
    const int G = (int)plan->workers.size();
    int home_idx = -1;
    GSB_TRY(plan_home(plan, mem, home_device, &home_idx));
    if (dim < 1 || dim > GSB_MAX_DIM) return fail(GSB_ERR_ARGUMENT, "dim must be in 1..8");
    if (!axis_len) return fail(GSB_ERR_ARGUMENT, "axis_len must not be NULL");
    if (n_batch < 1) return fail(GSB_ERR_ARGUMENT, "n_batch must be >= 1");
    int64_t n = 1, total_axes = 0;
    for (int t = 0; t < dim; ++t) {
        if (axis_len[t] < 0) return fail(GSB_ERR_ARGUMENT, "axis_len must be >= 0");
        n *= axis_len[t];
        total_axes += axis_len[t];
    }
    const int ncomp = vec ? dim : 1;
    const bool by_batch = n_batch >= G;      // ensembles: whole fields per device; else slabs of every field
    const int64_t len0 = axis_len[0];
    std::vector<int64_t> lens(axis_len, axis_len + dim);
    auto share_call = [=](int g, const double *c, const double *a, const double *b, const double *ax, int device,
                          void *st) -> int {
        int64_t lo, hi;
        const gsb_point_epilogue *pe = pepi ? pepi + g : nullptr;
        if (by_batch) {
            plan_share(n_batch, G, g, &lo, &hi);
            if (hi == lo) return GSB_OK;
            return structured_impl(c + lo * dim * n_modes, a + lo * n_modes, b + lo * n_modes, nullptr, ax, lens.data(),
                                   matrix, dim, n_modes, hi - lo, out + lo * ncomp * n, vec, epi, mem, device, st, pe);
        }
        plan_share(len0, G, g, &lo, &hi);
        if (hi == lo) return GSB_OK;
        const SlabOpt slab{lo, hi};
        return structured_impl(c, a, b, nullptr, ax, lens.data(), matrix, dim, n_modes, n_batch, out, vec, epi, mem, device,
                               st, pe, &slab);
    };
    if (mem == GSB_MEM_HOST) {
        std::vector<std::function<int()>> jobs((size_t)G);
        for (int g = 0; g < G; ++g) {
            const int device = plan->workers[(size_t)g]->device;
            jobs[(size_t)g] = [=]() { return share_call(g, cov, z1, z2, axes, device, nullptr); };
        }
        return plan_run(plan, jobs);
    }
    return plan_run_device(plan, home_idx, stream, [&](int g, cudaStream_t st, Scratch &scr) -> int {
        const bool local = (g == home_idx);
        const double *c, *a, *b, *ax;
        GSB_TRY(plan_stage(cov, (size_t)n_batch * dim * n_modes, local, &c, scr, st));
        GSB_TRY(plan_stage(z1, (size_t)n_batch * n_modes, local, &a, scr, st));
        GSB_TRY(plan_stage(z2, (size_t)n_batch * n_modes, local, &b, scr, st));
        GSB_TRY(plan_stage(axes, (size_t)total_axes, local, &ax, scr, st));
        return share_call(g, c, a, b, ax, plan->workers[(size_t)g]->device, st);
    });
}

static int plan_krige_impl(gsb_plan *plan, KrigeEvalArgs a, bool structured, int mem, int home_device, void *stream)
{
    if (!plan) return fail(GSB_ERR_ARGUMENT, "plan must not be NULL");
    std::lock_guard<std::mutex> lock(plan->call_mutex);
    const int G = (int)plan->workers.size();
    int home_idx = -1;
    GSB_TRY(plan_home(plan, mem, home_device, &home_idx));
    if (a.dim < 1 || a.dim > 4) return fail(GSB_ERR_ARGUMENT, "krige_evaluate: dim must be in 1..4");
    // points are independent, the kriging system is replicated: meshes are cut into slabs along axis 0 (contiguous
    // ranges of the C-ordered point index), flat point sets into contiguous ranges
    int64_t units = a.n, per_unit = 1, total_axes = 0;
    if (structured) {
        if (!a.axis_len) return fail(GSB_ERR_ARGUMENT, "axis_len must not be NULL");
        int64_t n = 1;
        for (int t = 0; t < a.dim; ++t) {
            if (a.axis_len[t] < 0) return fail(GSB_ERR_ARGUMENT, "axis_len must be >= 0");
            n *= a.axis_len[t];
            total_axes += a.axis_len[t];
        }
        units = a.axis_len[0];
        per_unit = units > 0 ? n / units : 0;
    }
    if (units < 0) return fail(GSB_ERR_ARGUMENT, "krige_evaluate: n_pts must be >= 0");
    auto share_call = [=](int g, KrigeEvalArgs args, int device, void *st) -> int {
        int64_t lo, hi;
        plan_share(units, G, g, &lo, &hi);
        if (hi == lo) return GSB_OK;
        args.first = lo * per_unit;
        args.count = (hi - lo) * per_unit;
        return krige_eval_impl(args, structured, mem, device, st);
    };
    if (mem == GSB_MEM_HOST) {
        std::vector<std::function<int()>> jobs((size_t)G);
        for (int g = 0; g < G; ++g) {
            const int device = plan->workers[(size_t)g]->device;
            jobs[(size_t)g] = [=]() { return share_call(g, a, device, nullptr); };
        }
        return plan_run(plan, jobs);
    }
    const int64_t n_tail = a.K - a.C - (a.unbiased ? 1 : 0);
    (void)n_tail;
    return plan_run_device(plan, home_idx, stream, [&](int g, cudaStream_t st, Scratch &scr) -> int {
        const bool local = (g == home_idx);
        KrigeEvalArgs args = a;
        // the system (K x K), the conditioning values / positions and the axes are re-read by every tile: private
        // copies; positions of a flat point set and drift rows are read once per point, straight from the home device
        if (a.K > 0) {
            GSB_TRY(plan_stage(a.mat, (size_t)a.K * a.K, local, &args.mat, scr, st));
            GSB_TRY(plan_stage(a.cond, (size_t)a.K, local, &args.cond, scr, st));
        }
        if (a.C > 0 && a.cond_pos) GSB_TRY(plan_stage(a.cond_pos, (size_t)a.dim * a.C, local, &args.cond_pos, scr, st));
        if (structured) GSB_TRY(plan_stage(a.axes, (size_t)total_axes, local, &args.axes, scr, st));
        return share_call(g, args, plan->workers[(size_t)g]->device, st);
    });
}

}  // namespace gsb

namespace gsb {

// both entry points: the chain itself (rng.py:77-101); `eval` is the log-pdf of a batch of radii
template <typename Eval>
static int sample_radii_impl(Eval &&eval, const uint32_t *mt_key_burn, int mt_pos_burn, const uint32_t *mt_key_main,
                             int mt_pos_main, const double *init, int nwalkers, int burn_in, int n_steps, double *chain)
{
    if (!mt_key_burn || !mt_key_main || mt_pos_burn < 0 || mt_pos_burn > 624 || mt_pos_main < 0 || mt_pos_main > 624 ||
        !init || !chain)
        return fail(GSB_ERR_ARGUMENT, "sample_radii_mcmc: NULL pointer or bad generator position");
    if (nwalkers < 2 || (nwalkers & 1) || burn_in < 0 || n_steps < 0)
        return fail(GSB_ERR_ARGUMENT, "sample_radii_mcmc: nwalkers must be even and >= 2, step counts >= 0");
    std::vector<double> coords(init, init + nwalkers), logp(nwalkers);
    for (int i = 0; i < nwalkers; ++i)
        if (!std::isfinite(coords[i])) return fail(GSB_ERR_ARGUMENT, "sample_radii_mcmc: initial guess not finite");
    if (eval(coords.data(), nwalkers, logp.data()))
        return fail(GSB_ERR_ARGUMENT, "sample_radii_mcmc: the log-pdf callback failed");
    for (int i = 0; i < nwalkers; ++i)
        if (std::isnan(logp[i])) return fail(GSB_ERR_ARGUMENT, "sample_radii_mcmc: initial log-pdf is NaN");
    Mt19937 rng;
    // each run_mcmc call of RNG.sample_ln_pdf gets its own generator state: rng.py:84-99 copies
    // self.random.get_state() into the sampler, and RNG.random is a fresh RandomState per access
    // (rng.py:193-203), seeded from the master generator
    for (int pass = 0; pass < 2; ++pass) {
        std::memcpy(rng.key, pass == 0 ? mt_key_burn : mt_key_main, sizeof rng.key);
        rng.pos = pass == 0 ? mt_pos_burn : mt_pos_main;
        const int rc = stretch_run(eval, rng, nwalkers, pass == 0 ? burn_in : n_steps, coords.data(), logp.data(),
                                   pass == 0 ? nullptr : chain);
        if (rc == 2) return fail(GSB_ERR_ARGUMENT, "sample_radii_mcmc: the log-pdf callback failed");
        if (rc) return fail(GSB_ERR_ARGUMENT, "sample_radii_mcmc: proposal or log-pdf not finite");
    }
    return GSB_OK;
}

}  // namespace gsb

// =============================================================================================
// C ABI
// =============================================================================================
using namespace gsb;

extern "C" {

int gsb_version(void) { return 100; }

const char *gsb_last_error(void) { return last_error_ref().c_str(); }

int gsb_device_count(int *count)
{
    if (!count) return fail(GSB_ERR_ARGUMENT, "count must not be NULL");
    int c = 0;
    if (cudaGetDeviceCount(&c) != cudaSuccess) {
        cudaGetLastError();
        c = 0;
    }
    *count = c;
    return GSB_OK;
}

int gsb_summate(const double *cov_samples, const double *z_1, const double *z_2, const double *pos,
                int64_t pos_ld, int dim, int64_t n_modes, int64_t n_pts, double *out, int mem,
                int device, void *stream)
{
    return summate_impl(cov_samples, z_1, z_2, nullptr, pos, pos_ld, dim, n_modes, n_pts, out, n_pts,
                        false, nullptr, mem, device, stream);
}

int gsb_summate_ex(const double *cov_samples, const double *z_1, const double *z_2, const double *pos,
                   int64_t pos_ld, int dim, int64_t n_modes, int64_t n_pts, double *out,
                   const gsb_epilogue *epi, int mem, int device, void *stream)
{
    return summate_impl(cov_samples, z_1, z_2, nullptr, pos, pos_ld, dim, n_modes, n_pts, out, n_pts,
                        false, epi, mem, device, stream);
}

int gsb_summate_incompr_ex(const double *cov_samples, const double *z_1, const double *z_2,
                           const double *pos, int64_t pos_ld, int dim, int64_t n_modes, int64_t n_pts,
                           double *out, int64_t out_ld, const gsb_epilogue *epi, int mem, int device,
                           void *stream)
{
    return summate_impl(cov_samples, z_1, z_2, nullptr, pos, pos_ld, dim, n_modes, n_pts, out, out_ld,
                        true, epi, mem, device, stream);
}

int gsb_summate_structured_ex(const double *cov_samples, const double *z_1, const double *z_2,
                              const double *axes, const int64_t *axis_len, const double *matrix, int dim,
                              int64_t n_modes, int64_t n_batch, double *out, const gsb_epilogue *epi,
                              int mem, int device, void *stream)
{
    return structured_impl(cov_samples, z_1, z_2, nullptr, axes, axis_len, matrix, dim, n_modes, n_batch,
                           out, false, epi, mem, device, stream);
}

int gsb_summate_incompr_structured_ex(const double *cov_samples, const double *z_1, const double *z_2,
                                      const double *axes, const int64_t *axis_len, const double *matrix,
                                      int dim, int64_t n_modes, int64_t n_batch, double *out,
                                      const gsb_epilogue *epi, int mem, int device, void *stream)
{
    return structured_impl(cov_samples, z_1, z_2, nullptr, axes, axis_len, matrix, dim, n_modes, n_batch,
                           out, true, epi, mem, device, stream);
}

int gsb_summate_incompr(const double *cov_samples, const double *z_1, const double *z_2,
                        const double *pos, int64_t pos_ld, int dim, int64_t n_modes, int64_t n_pts,
                        double *out, int64_t out_ld, int mem, int device, void *stream)
{
    return summate_impl(cov_samples, z_1, z_2, nullptr, pos, pos_ld, dim, n_modes, n_pts, out, out_ld,
                        true, nullptr, mem, device, stream);
}

int gsb_summate_structured(const double *cov_samples, const double *z_1, const double *z_2,
                           const double *axes, const int64_t *axis_len, const double *matrix, int dim,
                           int64_t n_modes, int64_t n_batch, double *out, int mem, int device,
                           void *stream)
{
    return structured_impl(cov_samples, z_1, z_2, nullptr, axes, axis_len, matrix, dim, n_modes, n_batch,
                           out, false, nullptr, mem, device, stream);
}

int gsb_summate_incompr_structured(const double *cov_samples, const double *z_1, const double *z_2,
                                   const double *axes, const int64_t *axis_len, const double *matrix,
                                   int dim, int64_t n_modes, int64_t n_batch, double *out, int mem,
                                   int device, void *stream)
{
    return structured_impl(cov_samples, z_1, z_2, nullptr, axes, axis_len, matrix, dim, n_modes, n_batch,
                           out, true, nullptr, mem, device, stream);
}

int gsb_summate_pp(const double *cov_samples, const double *z_1, const double *z_2, const double *pos,
                   int64_t pos_ld, int dim, int64_t n_modes, int64_t n_pts, double *out, const gsb_epilogue *epi,
                   const gsb_point_epilogue *pepi, int mem, int device, void *stream)
{
    return summate_impl(cov_samples, z_1, z_2, nullptr, pos, pos_ld, dim, n_modes, n_pts, out, n_pts, false, epi, mem,
                        device, stream, pepi);
}

int gsb_summate_structured_pp(const double *cov_samples, const double *z_1, const double *z_2, const double *axes,
                              const int64_t *axis_len, const double *matrix, int dim, int64_t n_modes,
                              int64_t n_batch, double *out, const gsb_epilogue *epi,
                              const gsb_point_epilogue *pepi, int mem, int device, void *stream)
{
    return structured_impl(cov_samples, z_1, z_2, nullptr, axes, axis_len, matrix, dim, n_modes, n_batch, out, false,
                           epi, mem, device, stream, pepi);
}

int gsb_cond_scaling(const double *error, int64_t n, double sill, double var, double *krige_var, double *gain,
                     int device, void *stream)
{
    if (n < 0) return fail(GSB_ERR_ARGUMENT, "n must be >= 0");
    if (n == 0) return GSB_OK;
    if (!error || !gain) return fail(GSB_ERR_ARGUMENT, "error and gain must not be NULL");
    DeviceGuard guard;
    DeviceState *dev = nullptr;
    GSB_TRY(ensure_device(device, &dev));
    const int blocks = (int)std::min<int64_t>((n + 255) / 256, 16LL * dev->sm_count);
    cond_scaling_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(error, n, sill, var, krige_var, gain);
    g_launches.fetch_add(1);
    GSB_CUDA(cudaGetLastError());
    return GSB_OK;
}

int gsb_summate_fourier(const double *spectrum_factor, const double *modes, const double *z_1,
                        const double *z_2, const double *pos, int64_t pos_ld, int dim, int64_t n_modes,
                        int64_t n_pts, double *out, int mem, int device, void *stream)
{
    if (n_modes > 0 && !spectrum_factor) return fail(GSB_ERR_ARGUMENT, "spectrum_factor must not be NULL");
    return summate_impl(modes, z_1, z_2, spectrum_factor, pos, pos_ld, dim, n_modes, n_pts, out, n_pts,
                        false, nullptr, mem, device, stream);
}

int gsb_summate_fourier_structured(const double *spectrum_factor, const double *modes, const double *z_1,
                                   const double *z_2, const double *axes, const int64_t *axis_len,
                                   const double *matrix, int dim, int64_t n_modes, double *out, int mem,
                                   int device, void *stream)
{
    if (n_modes > 0 && !spectrum_factor) return fail(GSB_ERR_ARGUMENT, "spectrum_factor must not be NULL");
    return structured_impl(modes, z_1, z_2, spectrum_factor, axes, axis_len, matrix, dim, n_modes, 1, out,
                           false, nullptr, mem, device, stream);
}

int gsb_calc_field_krige_and_variance(const double *krig_mat, const double *krig_vecs, int64_t vecs_ld,
                                      const double *cond, int64_t krige_size, int64_t n_pts, double *field,
                                      double *error, int mem, int device, void *stream)
{
    return krige_impl(krig_mat, krig_vecs, vecs_ld, cond, krige_size, n_pts, field, error, true, mem, device, stream);
}

int gsb_calc_field_krige(const double *krig_mat, const double *krig_vecs, int64_t vecs_ld, const double *cond,
                         int64_t krige_size, int64_t n_pts, double *field, int mem, int device, void *stream)
{
    return krige_impl(krig_mat, krig_vecs, vecs_ld, cond, krige_size, n_pts, field, nullptr, false, mem, device, stream);
}

int gsb_krige_evaluate(const gsb_cov_model *model, const double *krig_mat, const double *cond, int64_t krige_size,
                       const double *cond_pos, int64_t cond_no, int dim, const double *pos, int64_t pos_ld,
                       int64_t n_pts, int unbiased, const double *tail_rows, int64_t tail_ld, double *field,
                       double *error, int mem, int device, void *stream)
{
    KrigeEvalArgs a{model, krig_mat, cond, cond_pos, krige_size, cond_no, dim, pos, pos_ld, n_pts, nullptr, nullptr,
                    nullptr, unbiased, tail_rows, tail_ld, field, error};
    return krige_eval_impl(a, false, mem, device, stream);
}

int gsb_krige_evaluate_structured(const gsb_cov_model *model, const double *krig_mat, const double *cond,
                                  int64_t krige_size, const double *cond_pos, int64_t cond_no, int dim,
                                  const double *axes, const int64_t *axis_len, const double *matrix, int unbiased,
                                  const double *tail_rows, int64_t tail_ld, double *field, double *error, int mem,
                                  int device, void *stream)
{
    KrigeEvalArgs a{model, krig_mat, cond, cond_pos, krige_size, cond_no, dim, nullptr, 0, 0, axes, axis_len,
                    matrix, unbiased, tail_rows, tail_ld, field, error};
    return krige_eval_impl(a, true, mem, device, stream);
}

int gsb_sample_radii_mcmc(int pdf_kind, int dim, double len_rescaled, double nu, const uint32_t *mt_key_burn,
                          int mt_pos_burn, const uint32_t *mt_key_main, int mt_pos_main, const double *init,
                          int nwalkers, int burn_in, int n_steps, double *chain)
{
    if (pdf_kind != GSB_PDF_EXPONENTIAL && pdf_kind != GSB_PDF_MATERN && pdf_kind != GSB_PDF_GAUSSIAN)
        return fail(GSB_ERR_ARGUMENT, "sample_radii_mcmc: unknown pdf kind");
    if (dim < 1 || !(len_rescaled > 0.0) || (pdf_kind == GSB_PDF_MATERN && !(nu > 0.0)))
        return fail(GSB_ERR_ARGUMENT, "sample_radii_mcmc: bad model parameters");
    const RadPdf pdf(pdf_kind, dim, len_rescaled, nu);
    auto eval = [&pdf](const double *q, int n, double *out) -> int {
        for (int k = 0; k < n; ++k) out[k] = pdf.ln_pdf(q[k]);
        return 0;
    };
    return sample_radii_impl(eval, mt_key_burn, mt_pos_burn, mt_key_main, mt_pos_main, init, nwalkers, burn_in, n_steps,
                             chain);
}

int gsb_sample_radii_mcmc_cb(gsb_ln_pdf_fn ln_pdf, void *user, const uint32_t *mt_key_burn, int mt_pos_burn,
                             const uint32_t *mt_key_main, int mt_pos_main, const double *init, int nwalkers,
                             int burn_in, int n_steps, double *chain)
{
    if (!ln_pdf) return fail(GSB_ERR_ARGUMENT, "sample_radii_mcmc_cb: ln_pdf must not be NULL");
    auto eval = [ln_pdf, user](const double *q, int n, double *out) -> int { return ln_pdf(q, n, out, user); };
    return sample_radii_impl(eval, mt_key_burn, mt_pos_burn, mt_key_main, mt_pos_main, init, nwalkers, burn_in, n_steps,
                             chain);
}

int gsb_sample_modes_batch(int pdf_kind, int dim, double len_rescaled, double nu, const int64_t *seeds, int64_t n_seeds,
                           int64_t mode_no, int nwalkers, int burn_in, int64_t n_steps, double sample_around,
                           double two_pi, int n_threads, double *z_1, double *z_2, double *ang_1, double *ang_2,
                           double *rad)
{
    if (pdf_kind != GSB_PDF_EXPONENTIAL && pdf_kind != GSB_PDF_MATERN && pdf_kind != GSB_PDF_GAUSSIAN)
        return fail(GSB_ERR_ARGUMENT, "sample_modes_batch: unknown pdf kind");
    if (dim < 1 || dim > 3 || !(len_rescaled > 0.0) || (pdf_kind == GSB_PDF_MATERN && !(nu > 0.0)))
        return fail(GSB_ERR_ARGUMENT, "sample_modes_batch: bad model parameters (dim must be 1, 2 or 3)");
    if (n_seeds < 0 || mode_no < 1 || nwalkers < 2 || (nwalkers & 1) || burn_in < 0 || n_steps < 1 ||
        n_steps * (int64_t)nwalkers > (int64_t)0x7fffffff)
        return fail(GSB_ERR_ARGUMENT, "sample_modes_batch: bad sizes");
    if (n_seeds > 0 && (!seeds || !z_1 || !z_2 || !ang_1 || !rad || (dim == 3 && !ang_2)))
        return fail(GSB_ERR_ARGUMENT, "sample_modes_batch: NULL pointer");
    for (int64_t i = 0; i < n_seeds; ++i)
        if (seeds[i] < 0 || seeds[i] > (int64_t)0xffffffffLL)
            return fail(GSB_ERR_ARGUMENT, "sample_modes_batch: seeds must fit 32 bits (numpy legacy seeding)");
    ModeBatchArgs a{RadPdf(pdf_kind, dim, len_rescaled, nu), dim, mode_no, nwalkers, burn_in, n_steps, sample_around, two_pi};
    const int nt = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(n_threads > 0 ? n_threads : 1, n_seeds), 256));
    std::atomic<int64_t> next{0};
    std::atomic<int> bad{0};
    auto work = [&]() {
        std::vector<double> chain;
        for (;;) {
            const int64_t i = next.fetch_add(1);
            if (i >= n_seeds) break;
            if (sample_modes_one(a, (uint32_t)seeds[i], z_1 + i * mode_no, z_2 + i * mode_no, ang_1 + i * mode_no,
                                 dim == 3 ? ang_2 + i * mode_no : nullptr, rad + i * mode_no, chain))
                bad.store(1);
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; ++t) pool.emplace_back(work);
    work();
    for (auto &th : pool) th.join();
    if (bad.load()) return fail(GSB_ERR_ARGUMENT, "sample_modes_batch: proposal or log-pdf not finite");
    return GSB_OK;
}

int gsb_scale_shift(double *field, int64_t n, double scale, double shift, int device, void *stream)
{
    if (n < 0) return fail(GSB_ERR_ARGUMENT, "n must be >= 0");
    if (n == 0) return GSB_OK;
    if (!field) return fail(GSB_ERR_ARGUMENT, "field must not be NULL");
    DeviceGuard guard;
    DeviceState *dev = nullptr;
    GSB_TRY(ensure_device(device, &dev));
    const int blocks = (int)std::min<int64_t>((n + 255) / 256, 16LL * dev->sm_count);
    scale_shift_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(field, n, scale, shift);
    g_launches.fetch_add(1);
    GSB_CUDA(cudaGetLastError());
    return GSB_OK;
}

int gsb_plan_krige_evaluate(gsb_plan *plan, const gsb_cov_model *model, const double *krig_mat, const double *cond,
                            int64_t krige_size, const double *cond_pos, int64_t cond_no, int dim, const double *pos,
                            int64_t pos_ld, int64_t n_pts, const double *axes, const int64_t *axis_len,
                            const double *matrix, int unbiased, const double *tail_rows, int64_t tail_ld, double *field,
                            double *error, int mem, int home_device, void *stream)
{
    const bool structured = (pos == nullptr);
    KrigeEvalArgs a{model, krig_mat, cond, cond_pos, krige_size, cond_no, dim, pos, pos_ld, n_pts, axes, axis_len,
                    matrix, unbiased, tail_rows, tail_ld, field, error};
    return plan_krige_impl(plan, a, structured, mem, home_device, stream);
}

int gsb_summate_structured_slab(const double *cov_samples, const double *z_1, const double *z_2, const double *axes,
                                const int64_t *axis_len, const double *matrix, int dim, int64_t n_modes,
                                int64_t n_batch, int64_t slab_lo, int64_t slab_hi, double *out, int incompr,
                                const gsb_epilogue *epi, const gsb_point_epilogue *pepi, int mem, int device,
                                void *stream)
{
    const SlabOpt slab{slab_lo, slab_hi};
    return structured_impl(cov_samples, z_1, z_2, nullptr, axes, axis_len, matrix, dim, n_modes, n_batch, out,
                           incompr != 0, epi, mem, device, stream, pepi, &slab);
}

int gsb_ipc_export(const void *dev_ptr, int device, unsigned char *handle, int64_t *offset)
{
    if (!dev_ptr || !handle || !offset) return fail(GSB_ERR_ARGUMENT, "ipc_export: NULL pointer");
    static_assert(sizeof(cudaIpcMemHandle_t) == GSB_IPC_HANDLE_BYTES, "handle size");
    DeviceGuard guard;
    DeviceState *dev = nullptr;
    GSB_TRY(ensure_device(device, &dev));
    cudaIpcMemHandle_t h;
    GSB_CUDA(cudaIpcGetMemHandle(&h, const_cast<void *>(dev_ptr)));
    // the handle names the whole allocation: report where dev_ptr lies inside it
    // (driver entry point through the runtime: the library does not link libcuda, so it still loads on a machine
    // without a driver -- where every compute entry fails loudly instead)
    typedef CUresult (*range_fn)(CUdeviceptr *, size_t *, CUdeviceptr);
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    GSB_CUDA(cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qres));
    CUdeviceptr base = 0;
    size_t size = 0;
    if (!fn || qres != cudaDriverEntryPointSuccess ||
        reinterpret_cast<range_fn>(fn)(&base, &size, reinterpret_cast<CUdeviceptr>(dev_ptr)) != CUDA_SUCCESS)
        return fail(GSB_ERR_CUDA, "ipc_export: cuMemGetAddressRange failed");
    std::memcpy(handle, &h, sizeof h);
    *offset = (int64_t)(reinterpret_cast<CUdeviceptr>(dev_ptr) - base);
    return GSB_OK;
}

int gsb_ipc_open(const unsigned char *handle, int device, void **base)
{
    if (!handle || !base) return fail(GSB_ERR_ARGUMENT, "ipc_open: NULL pointer");
    DeviceGuard guard;
    DeviceState *dev = nullptr;
    GSB_TRY(ensure_device(device, &dev));
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, sizeof h);
    GSB_CUDA(cudaIpcOpenMemHandle(base, h, cudaIpcMemLazyEnablePeerAccess));
    return GSB_OK;
}

int gsb_ipc_close(void *base, int device)
{
    if (!base) return GSB_OK;
    DeviceGuard guard;
    DeviceState *dev = nullptr;
    GSB_TRY(ensure_device(device, &dev));
    GSB_CUDA(cudaIpcCloseMemHandle(base));
    return GSB_OK;
}

int gsb_plan_share(int64_t n, int parts, int part, int64_t *lo, int64_t *hi)
{
    if (!lo || !hi) return fail(GSB_ERR_ARGUMENT, "NULL output pointer");
    if (n < 0 || parts < 1 || part < 0 || part >= parts) return fail(GSB_ERR_ARGUMENT, "plan_share: invalid part / parts");
    plan_share(n, parts, part, lo, hi);
    return GSB_OK;
}

int gsb_plan_create(const int *devices, int n_devices, gsb_plan **plan)
{
    if (!plan) return fail(GSB_ERR_ARGUMENT, "plan must not be NULL");
    *plan = nullptr;
    DeviceGuard guard;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        cudaGetLastError();
        return fail(GSB_ERR_NO_DEVICE, "no CUDA device available: the B200 backend has no CPU fallback");
    }
    std::vector<int> devs;
    if (!devices || n_devices <= 0) {
        for (int d = 0; d < count && d < 64; ++d) devs.push_back(d);
    } else {
        devs.assign(devices, devices + n_devices);
    }
    // (a device may be listed more than once: its shares then run one after the other -- this is how the share
    // logic is tested on a one-GPU box)
    if (devs.size() > 64) return fail(GSB_ERR_ARGUMENT, "plan: too many devices");
    for (size_t i = 0; i < devs.size(); ++i)
        if (devs[i] < 0 || devs[i] >= count || devs[i] >= 64) return fail(GSB_ERR_ARGUMENT, "plan: invalid device index");
    std::unique_ptr<gsb_plan> p(new gsb_plan);
    for (int d : devs) {
        DeviceState *dev = nullptr;
        GSB_TRY(ensure_device(d, &dev));        // streams, pool settings; leaves d current
        std::unique_ptr<PlanWorker> w(new PlanWorker);
        w->device = d;
        GSB_CUDA(cudaEventCreateWithFlags(&w->begin_ev, cudaEventDisableTiming));
        GSB_CUDA(cudaEventCreateWithFlags(&w->done_ev, cudaEventDisableTiming));
        // peer access to every other device of the plan (the device route stores into the home device's memory)
        for (int o : devs) {
            if (o == d) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, d, o) != cudaSuccess || !can) {
                cudaGetLastError();
                p->peer_ok = false;
                continue;
            }
            const cudaError_t e = cudaDeviceEnablePeerAccess(o, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) p->peer_ok = false;
            cudaGetLastError();
        }
        p->workers.push_back(std::move(w));
    }
    for (auto &w : p->workers) {
        PlanWorker *raw = w.get();
        raw->th = std::thread([raw]() { raw->loop(); });
    }
    *plan = p.release();
    return GSB_OK;
}

int gsb_plan_destroy(gsb_plan *plan)
{
    if (!plan) return GSB_OK;
    DeviceGuard guard;
    {
        std::lock_guard<std::mutex> lock(plan->call_mutex);
        for (auto &w : plan->workers) {
            {
                std::lock_guard<std::mutex> lk(w->m);
                w->quit = true;
                w->cv.notify_all();
            }
            if (w->th.joinable()) w->th.join();
            if (cudaSetDevice(w->device) == cudaSuccess) {
                cudaStreamSynchronize(g_dev[w->device].streams[2]);
                cudaEventDestroy(w->begin_ev);
                cudaEventDestroy(w->done_ev);
            }
            cudaGetLastError();
        }
    }
    delete plan;
    return GSB_OK;
}

int gsb_plan_info(const gsb_plan *plan, int *n_devices, int *devices, int *peer_access)
{
    if (!plan || !n_devices) return fail(GSB_ERR_ARGUMENT, "plan and n_devices must not be NULL");
    *n_devices = (int)plan->workers.size();
    if (devices)
        for (size_t g = 0; g < plan->workers.size(); ++g) devices[g] = plan->workers[g]->device;
    if (peer_access) *peer_access = plan->peer_ok ? 1 : 0;
    return GSB_OK;
}

int gsb_plan_summate(gsb_plan *plan, const double *cov_samples, const double *z_1, const double *z_2, const double *pos,
                     int64_t pos_ld, int dim, int64_t n_modes, int64_t n_pts, double *out, int64_t out_ld, int incompr,
                     const gsb_epilogue *epi, const gsb_point_epilogue *pepi, int mem, int home_device, void *stream)
{
    return plan_summate_impl(plan, cov_samples, z_1, z_2, pos, pos_ld, dim, n_modes, n_pts, out, out_ld, incompr != 0, epi,
                             pepi, mem, home_device, stream);
}

int gsb_plan_summate_structured(gsb_plan *plan, const double *cov_samples, const double *z_1, const double *z_2,
                                const double *axes, const int64_t *axis_len, const double *matrix, int dim,
                                int64_t n_modes, int64_t n_batch, double *out, int incompr, const gsb_epilogue *epi,
                                const gsb_point_epilogue *pepi, int mem, int home_device, void *stream)
{
    return plan_structured_impl(plan, cov_samples, z_1, z_2, axes, axis_len, matrix, dim, n_modes, n_batch, out,
                                incompr != 0, epi, pepi, mem, home_device, stream);
}

int gsb_release_memory(int device)
{
    DeviceGuard guard;
    DeviceState *dev = nullptr;
    GSB_TRY(ensure_device(device, &dev));
    std::lock_guard<std::mutex> lock(dev->call_mutex);
    GSB_CUDA(cudaDeviceSynchronize());
    cudaMemPool_t pool;
    GSB_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
    GSB_CUDA(cudaMemPoolTrimTo(pool, 0));
    return GSB_OK;
}

int gsb_set_option(const char *name, int64_t value)
{
    if (!name) return fail(GSB_ERR_ARGUMENT, "option name must not be NULL");
    const std::string n(name);
    if (n == "structured_min_tiles") g_opt_structured_min_tiles = value;
    else if (n == "force_path") g_opt_force_path = value;
    else if (n == "host_chunk_points") g_opt_host_chunk_points = value;
    else if (n == "scratch_mb") g_opt_scratch_mb = std::max<int64_t>(value, 1);
    else if (n == "min_chunks" || n == "sep_path" || n == "chunk_growth_pct" || n == "partial_tiles") {
        // tuning knobs of the first-generation separable path (removed): accepted and ignored
    }
    else if (n == "trace") g_opt_trace = value;
    else if (n == "direct_cfg") g_opt_direct_cfg = value;
    else if (n == "direct_split") g_opt_direct_split = value;
    else if (n == "time_kernels") g_opt_time_kernels = value;
    else if (n == "fold_axes") g_opt_fold_axes = value;
    else if (n == "krige_host_chunk_mb") g_opt_krige_host_chunk_mb = std::max<int64_t>(value, 1);
    else if (n == "sk_table_mb") g_opt_sk_table_mb = std::max<int64_t>(value, 1);
    else if (n == "sk_grid") g_opt_sk_grid = std::max<int64_t>(value, 0);
    else if (n == "host_pieces") g_opt_host_pieces = std::max<int64_t>(value, 0);
    else if (n == "sk_pack") g_opt_sk_pack = std::max<int64_t>(value, 0);
    else return fail(GSB_ERR_ARGUMENT, "unknown option: " + n);
    return GSB_OK;
}

int64_t gsb_get_counter(const char *name)
{
    if (!name) return -1;
    const std::string n(name);
    if (n == "launches") return g_launches.load();
    if (n == "direct_calls") return g_cnt_direct.load();
    if (n == "separable_calls") return g_cnt_separable.load();
    if (n == "krige_calls") return g_cnt_krige.load();
    if (n == "folded_calls") return g_cnt_folded.load();
    if (n == "sk_calls") return g_cnt_sk.load();
    if (n == "packed_calls") return g_cnt_packed.load();
    return -1;
}

int gsb_streamk_plan(int64_t tile_begin, int64_t tile_end, int64_t ly, int64_t lc, int n_stages, int max_grid,
                     int64_t *tiles, int32_t *stages, int *grid)
{
    if (!tiles || !stages || !grid) return fail(GSB_ERR_ARGUMENT, "NULL output pointer");
    if (tile_begin < 0 || tile_end < tile_begin || ly < 1 || lc < 1 || n_stages < 1 || max_grid < 1)
        return fail(GSB_ERR_ARGUMENT, "streamk_plan: bad geometry");
    std::vector<SkBound> bnd;
    const int n_yt = (int)((ly + SK_TM - 1) / SK_TM), n_ct = (int)((lc + SK_TN - 1) / SK_TN);
    const int g = sk_plan(tile_begin, tile_end, n_yt, n_ct, n_stages, ly, lc, max_grid, bnd);
    for (int c = 0; c <= g; ++c) {
        tiles[c] = bnd[(size_t)c].tile;
        stages[c] = bnd[(size_t)c].stage;
    }
    *grid = g;
    return GSB_OK;
}

int gsb_kernel_times(double *total_ms, int64_t *n_launches)
{
    if (!total_ms || !n_launches) return fail(GSB_ERR_ARGUMENT, "NULL output pointer");
    std::lock_guard<std::mutex> lock(g_time_mutex);
    double tot = 0.0;
    int64_t n = 0;
    for (auto &ev : g_time_events) {
        float ms = 0.f;
        GSB_CUDA(cudaEventSynchronize(ev.second));
        GSB_CUDA(cudaEventElapsedTime(&ms, ev.first, ev.second));
        tot += ms;
        ++n;
        cudaEventDestroy(ev.first);
        cudaEventDestroy(ev.second);
    }
    g_time_events.clear();
    *total_ms = tot;
    *n_launches = n;
    return GSB_OK;
}

int gsb_measure_fp64_peak(int device, int kind, double seconds, double *fma_per_s)
{
    if (!fma_per_s) return fail(GSB_ERR_ARGUMENT, "fma_per_s must not be NULL");
    DeviceGuard guard;
    DeviceState *dev = nullptr;
    GSB_TRY(ensure_device(device, &dev));
    cudaStream_t st = dev->streams[0];
    double *sink = nullptr;
    GSB_CUDA(cudaMalloc(&sink, sizeof(double)));
    cudaEvent_t e0, e1;
    GSB_CUDA(cudaEventCreate(&e0));
    GSB_CUDA(cudaEventCreate(&e1));
    const int blocks = dev->sm_count * 8, threads = 256, iters = 2000;
    // FMAs per launch
    const double per_launch = (kind == 0)
        ? (double)blocks * threads * iters * 16.0 * 8.0
        : (double)blocks * (threads / 32) * iters * 4.0 * 8.0 * 256.0;  // m8n8k4 = 256 FMA per warp
    auto launch = [&]() {
        if (kind == 0) dfma_peak_kernel<<<blocks, threads, 0, st>>>(sink, iters, 1.0);
        else dmma_peak_kernel<<<blocks, threads, 0, st>>>(sink, iters, 1.0);
        g_launches.fetch_add(1);
    };
    for (int w = 0; w < 3; ++w) launch();
    GSB_CUDA(cudaStreamSynchronize(st));
    double best = 0.0;
    const auto t_end = std::chrono::steady_clock::now() + std::chrono::duration<double>(std::max(0.05, seconds));
    do {
        GSB_CUDA(cudaEventRecord(e0, st));
        for (int r = 0; r < 5; ++r) launch();
        GSB_CUDA(cudaEventRecord(e1, st));
        GSB_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        GSB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        best = std::max(best, 5.0 * per_launch / (ms * 1e-3));
    } while (std::chrono::steady_clock::now() < t_end);
    GSB_CUDA(cudaGetLastError());
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    *fma_per_s = best;
    return GSB_OK;
}

}  // extern "C"
