#!/usr/bin/env python
"""Benchmark of the hot path: randomisation-method field summation (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload c2]

(`--workload c1|c3|c4|c5` run the other BASELINE configs, `--workload krige` the widened row f1: the
kriging evaluation of config 5 in points/s, same JSON contract.)

Workload (default, the config BASELINE.json's metric is quoted on): configs[1] -- SRF Exponential
3D on a structured 512^3 mesh (134 217 728 points), mode_no = 1000; the mode set is the one the
unmodified reference draws for seed 20170519 (tests/golden/config_modes.npz).  A "step" is one
evaluation of the whole field.  With N > 1 GPUs the mesh is sharded into slabs along axis 0, one
process per GPU, no collective on the data path (strong scaling: the total work is fixed, as
BASELINE.json configs[1] states: "points sharded over 1/2/4/8 GPUs").

Printed JSON (one line, rank 0):
  value     point*modes/s, device-resident inputs, CUDA-event time summed over exactly K steps,
            max over ranks (L2 is flushed between steps outside the event brackets)
  e2e       same metric through the public numpy API gstools_b200.summate_structured():
            host buffers in, host field out; H2D + D2H inside the timed region
  roofline  FP64 pipe: algorithmic work = 2 DFMA (4 flop) per (point, mode) for the separable
            kernel; peak = DFMA rate measured live by gsb_measure_fp64_peak
  cpu_baseline  the C/OpenMP oracle (a port of the reference loop nest; the reference's own
            gstools-cython / gstools_core binaries are not available) on a bounded point sample
--impl reference times that CPU implementation as the whole step.
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

import bench_configs as bc  # noqa: E402

METRIC = "RandMeth point*modes/s (fp64)"
UNIT = "point*modes/s"
NOMINAL_DFMA_PER_S = 148 * 64 * 1.965e9  # 18.61e12, SURVEY.md 8(d)


# ----------------------------------------------------------------------------------------------
# workloads
# ----------------------------------------------------------------------------------------------
def make_workload(tag):
    if tag == "c2":
        cfg = bc.config2(512)
        cfg.update(path="separable", work_per_pair_dfma=2, tag="C2")
    elif tag == "c4":
        cfg = bc.config4(256)
        cfg.update(path="separable_incompr", work_per_pair_dfma=6, tag="C4")
    elif tag == "c5":
        cfg = bc.config5(128, 256)
        cfg.update(path="separable_batched", work_per_pair_dfma=2, tag="C5")
    elif tag == "c3":
        cfg = bc.config3(20_000_000)
        # executed FP64-pipe instructions per pair (DFMA 9 + DMUL 3 + DADD 3, verified in the ncu opcode
        # histogram); SURVEY.md 8(d)'s contract budget is d + 2 + 20 = 24 (libdevice-class sincos)
        cfg.update(path="direct", work_per_pair_dfma=2 + 13, contract_per_pair_dfma=2 + 2 + 20, tag="C3")
    elif tag == "c1":
        cfg = bc.config1()
        cfg.update(path="separable", work_per_pair_dfma=2, tag="C1")
    else:
        raise SystemExit(f"unknown workload {tag}")
    return cfg


def pairs_of(cfg):
    n_modes = cfg["cov"].shape[-1]
    n_batch = cfg["cov"].shape[0] if cfg["cov"].ndim == 3 else 1
    if "axes" in cfg and cfg["path"] != "direct":
        n = int(np.prod([len(a) for a in cfg["axes"]]))
    else:
        n = cfg["pos"].shape[1]
    return n * n_modes * n_batch


# ----------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock, power and throttle reasons DURING the timed region (NVML, 20 ms period;
    falls back to polling nvidia-smi)."""

    def __init__(self, index):
        self.index = index
        self.samples = []     # (sm_mhz, sm_max_mhz, power_w, reasons_bitmask or list)
        self._stop = threading.Event()
        self._thr = None
        self._nvml = None
        try:
            import pynvml

            pynvml.nvmlInit()
            uuid_index = index
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                try:
                    uuid_index = int(vis.split(",")[index])
                except ValueError:
                    uuid_index = index
            self._h = pynvml.nvmlDeviceGetHandleByIndex(uuid_index)
            self._nvml = pynvml
        except Exception:
            self._nvml = None

    def _sample_nvml(self):
        n = self._nvml
        sm = n.nvmlDeviceGetClockInfo(self._h, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(self._h, n.NVML_CLOCK_SM)
        pw = n.nvmlDeviceGetPowerUsage(self._h) / 1000.0
        mask = n.nvmlDeviceGetCurrentClocksEventReasons(self._h)
        reasons = []
        for name, attr in (("hw_slowdown", "nvmlClocksEventReasonHwSlowdown"),
                           ("hw_thermal_slowdown", "nvmlClocksEventReasonHwThermalSlowdown"),
                           ("sw_thermal_slowdown", "nvmlClocksEventReasonSwThermalSlowdown"),
                           ("sw_power_cap", "nvmlClocksEventReasonSwPowerCap")):
            bit = getattr(n, attr, None)
            if bit is None:
                bit = getattr(n, attr.replace("ClocksEventReason", "ClocksThrottleReason"), 0)
            if mask & bit:
                reasons.append(name)
        self.samples.append((float(sm), float(mx), float(pw), reasons))

    def _sample_smi(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                              "--format=csv,noheader,nounits"], capture_output=True, text=True,
                             timeout=5).stdout
        p = [x.strip() for x in out.strip().split(",")]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for i, nm in enumerate(names) if p[3 + i].lower().startswith("active")]
        self.samples.append((float(p[0]), float(p[1]), float(p[2]), reasons))

    def _run(self):
        while not self._stop.is_set():
            try:
                if self._nvml is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            self._stop.wait(0.02)

    def __enter__(self):
        self._thr = threading.Thread(target=self._run, daemon=True)
        self._thr.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._thr.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        reasons = sorted({r for s in self.samples for r in s[3]})
        return {"sm_mhz": statistics.median(s[0] for s in self.samples),
                "sm_min_mhz": min(s[0] for s in self.samples),
                "sm_max_mhz": max(s[1] for s in self.samples),
                "power_w_max": max(s[2] for s in self.samples), "reasons": reasons,
                "samples": len(self.samples),
                "source": "nvml" if self._nvml is not None else "nvidia-smi"}


# ----------------------------------------------------------------------------------------------
# CPU legs (oracle/ is test + baseline infrastructure; never the product path)
# ----------------------------------------------------------------------------------------------
def cpu_sample_points(cfg, n_pts, offset=0):
    """A contiguous range of the workload's points as a flat (dim, m) array."""
    if cfg["path"] == "direct":
        return np.ascontiguousarray(cfg["pos"][:, offset:offset + n_pts])
    n = int(np.prod([len(a) for a in cfg["axes"]]))
    idx = (np.arange(n_pts) + offset) % n
    return bc.grid_points(cfg["axes"], cfg.get("matrix"), idx)


def host_threads():
    """Threads of the CPU arm: every core this process may run on, set EXPLICITLY.  The reference uses all
    cores (src/gstools/config.py:8, NUM_THREADS = None -> all); torchrun exports OMP_NUM_THREADS=1, which
    must not silently turn the N > 1 reference arm into a single-threaded run."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_run(cfg, pos):
    import oracle

    cov, z1, z2 = cfg["cov"], cfg["z1"], cfg["z2"]
    if cov.ndim == 3:
        cov, z1, z2 = cov[0], z1[0], z2[0]
    fn = oracle.summate_incompr if cfg["kind"] == "incompr" else oracle.summate
    t0 = time.perf_counter()
    fn(cov, z1, z2, pos, num_threads=host_threads())
    return time.perf_counter() - t0


def cpu_baseline(cfg, target_seconds=12.0):
    import oracle

    oracle.build()
    n_modes = cfg["cov"].shape[-1]
    probe = 65536
    t = cpu_run(cfg, cpu_sample_points(cfg, probe))
    rate = probe * n_modes / max(t, 1e-9)
    m = int(min(max(probe, rate * target_seconds / n_modes), 16_000_000))
    t = cpu_run(cfg, cpu_sample_points(cfg, m))
    return {"value": m * n_modes / t, "unit": UNIT, "cores": host_threads(), "kind": "port",
            "sample": f"{m} contiguous points of the workload x {n_modes} modes "
                      f"({m * n_modes:.3g} pairs, {t:.1f} s, C/OpenMP oracle, glibc libm)"}


def run_reference_arm(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle

    oracle.build()
    n_modes = cfg["cov"].shape[-1]
    probe = 65536
    t = cpu_run(cfg, cpu_sample_points(cfg, probe))
    rate = probe * n_modes / max(t, 1e-9)
    budget = 150.0 / max(1, args.steps + args.warmup)          # whole run within a few minutes
    m = int(min(max(probe, rate * min(budget, 20.0) / n_modes), 4_000_000))
    for w in range(args.warmup):
        cpu_run(cfg, cpu_sample_points(cfg, m, offset=w * m))
    times = [cpu_run(cfg, cpu_sample_points(cfg, m, offset=(args.warmup + k) * m)) for k in range(args.steps)]
    total = sum(times)
    value = args.steps * m * n_modes / total
    sample = (f"each step = {m} contiguous points of the workload x {n_modes} modes on the host CPU "
              f"(C/OpenMP port of the reference loop nest, glibc libm); full-size time extrapolates linearly")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(cfg, args),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": host_threads(), "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(cfg, args):
    c = {"workload": cfg["name"], "baseline_config": cfg["tag"], "path": cfg["path"],
         "mode_no": int(cfg["cov"].shape[-1]), "dim": int(cfg["dim"]),
         "modes": "drawn by the unmodified reference (seed 20170519), tests/golden/config_modes.npz",
         "sharding": f"axis-0 slabs over {args.gpus} GPU(s), no collective",
         "l2": "256 MiB flush buffer written between timed steps, outside the event brackets; "
               "the 1.07 GB output exceeds L2"}
    return c


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
def run_b200_arm(args, cfg):
    import torch
    import torch.distributed as dist

    import gstools_b200 as gsb
    from gstools_b200.dist import shard_range

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 backend has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    gsb.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- this rank's share of the workload ----
    structured = cfg["path"] != "direct"
    vec = cfg["kind"] == "incompr"
    batched = cfg["cov"].ndim == 3
    if batched:                       # ensembles shard over realisations
        lo, hi = shard_range(cfg["cov"].shape[0], rank, world)
        h_cov, h_z1, h_z2 = cfg["cov"][lo:hi], cfg["z1"][lo:hi], cfg["z2"][lo:hi]
        h_axes = [np.ascontiguousarray(a) for a in cfg["axes"]]
    elif structured:                  # meshes shard over axis-0 slabs
        lo, hi = shard_range(len(cfg["axes"][0]), rank, world)
        h_cov, h_z1, h_z2 = cfg["cov"], cfg["z1"], cfg["z2"]
        h_axes = [np.ascontiguousarray(cfg["axes"][0][lo:hi])] + [np.ascontiguousarray(a) for a in cfg["axes"][1:]]
    else:                             # point sets shard over contiguous ranges
        lo, hi = shard_range(cfg["pos"].shape[1], rank, world)
        h_cov, h_z1, h_z2 = cfg["cov"], cfg["z1"], cfg["z2"]
        h_pos = np.ascontiguousarray(cfg["pos"][:, lo:hi])
    total_pairs = pairs_of(cfg)

    def pin(a):
        return torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()

    h_cov, h_z1, h_z2 = pin(h_cov), pin(h_z1), pin(h_z2)
    d_cov, d_z1, d_z2 = (torch.tensor(a, device=dev) for a in (h_cov, h_z1, h_z2))
    # --plan G: ONE process drives G GPUs (gsb_plan_*): the library cuts the call into slabs / point ranges /
    # batch entries itself; device route = every GPU stores its share into the result tensor on GPU 0 (compute and
    # gather are one kernel), host route = every GPU copies its share into its slice of the one pinned array
    plan = None
    if args.plan > 1:
        if world != 1:
            raise SystemExit("--plan is the single-process mode: do not combine it with torchrun")
        plan = gsb.Plan(list(range(args.plan)))
    if structured:
        h_axes = [pin(a) for a in h_axes]
        d_axes = [torch.tensor(a, device=dev) for a in h_axes]
        fn = gsb.summate_incompr_structured if vec else gsb.summate_structured
        if plan is not None:
            fn = plan.summate_incompr_structured if vec else plan.summate_structured

        def step_device():
            return fn(d_cov, d_z1, d_z2, d_axes, cfg.get("matrix"))

        def step_host():
            return fn(h_cov, h_z1, h_z2, h_axes, cfg.get("matrix"))

        h2d = h_cov.nbytes + h_z1.nbytes + h_z2.nbytes + sum(a.nbytes for a in h_axes)
    else:
        h_pos = pin(h_pos)
        d_pos = torch.tensor(h_pos, device=dev)
        fn = gsb.summate_incompr if vec else gsb.summate
        if plan is not None:
            fn = plan.summate_incompr if vec else plan.summate

        def step_device():
            return fn(d_cov, d_z1, d_z2, d_pos)

        def step_host():
            return fn(h_cov, h_z1, h_z2, h_pos)

        h2d = h_cov.nbytes + h_z1.nbytes + h_z2.nbytes + h_pos.nbytes

    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    peak_dfma = gsb.measure_fp64_peak(local, 0, 0.4)
    peak_dmma = gsb.measure_fp64_peak(local, 1, 0.3)

    # ---- device-resident throughput ----
    for _ in range(args.warmup):
        out = step_device()
        del out
    barrier()
    gsb.set_option("time_kernels", 1)
    gsb.kernel_times()
    launches0 = gsb.get_counter("launches")
    pairs = []
    with ClockSampler(local) as clocks:
        t_wall0 = time.perf_counter()
        for _ in range(args.steps):
            flush.fill_(1)
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            out = step_device()
            e1.record()
            pairs.append((e0, e1))
            del out
        barrier()
        t_wall = time.perf_counter() - t_wall0
    launches = gsb.get_counter("launches") - launches0
    kern_ms, kern_n = gsb.kernel_times()
    gsb.set_option("time_kernels", 0)
    dev_s = sum(a.elapsed_time(b) for a, b in pairs) * 1e-3
    t = torch.tensor([dev_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_s = float(t.item())
    value = args.steps * total_pairs / dev_s

    # ---- --gather: the same step delivered as ONE device array on rank 0 (sum + gather) ----
    gather = None
    gather_mode = args.gather
    if gather_mode == "auto":     # default: measure the p2p gather next to the plain step, leave `value` alone
        gather_mode = "p2p" if (world > 1 and structured and not batched) else "none"
    if gather_mode != "none" and world > 1 and structured and not batched:
        from gstools_b200 import dist as gdist

        g_axes = [torch.tensor(np.ascontiguousarray(a), device=dev) for a in cfg["axes"]]

        def step_gather():
            return gdist.summate_structured_gathered(d_cov, d_z1, d_z2, g_axes, cfg.get("matrix"), dst=0,
                                                     incompr=vec, mode=gather_mode, pieces=args.pieces,
                                                     reserve_sms=args.reserve_sms)

        try:
            for _ in range(args.warmup):
                out = step_gather()
                del out
        except RuntimeError as exc:       # raised on every rank together (open_peer_field): fall back, do not hang
            if rank == 0:
                print(f"bench.py: p2p gather unavailable ({exc}); using the NCCL gather", file=sys.stderr)
            gather_mode = "nccl"
            for _ in range(args.warmup):
                out = step_gather()
                del out
        barrier()
        gpairs = []
        for _ in range(args.steps):
            flush.fill_(1)
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            out = step_gather()
            e1.record()
            gpairs.append((e0, e1))
            del out
        barrier()
        g_s = sum(a.elapsed_time(b) for a, b in gpairs) * 1e-3
        t = torch.tensor([g_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        g_s = float(t.item())
        n_full = int(np.prod([len(a) for a in cfg["axes"]])) * (cfg["dim"] if vec else 1)
        gather = {"mode": gather_mode, "pieces": args.pieces if gather_mode == "nccl" else 1,
                  "reserve_sms": args.reserve_sms if gather_mode == "nccl" else 0,
                  "ms_per_step": 1e3 * g_s / args.steps, "compute_only_ms_per_step": 1e3 * dev_s / args.steps,
                  "exposed_ms": 1e3 * (g_s - dev_s) / args.steps,
                  "bytes_into_rank0_per_step": int(8 * n_full * (world - 1) / world),
                  "value_no_gather": value,
                  "what": "rank 0 ends the step holding the whole field as one CUDA tensor; "
                          + ("row pieces sent with grouped NCCL send/recv on a side stream while the next piece contracts"
                             if gather_mode == "nccl" else
                             "every rank's contraction kernel stores its slab into rank 0's tensor over NVLink (CUDA IPC "
                             "mapping), a one-element all-reduce orders the step")}
        gather["value"] = args.steps * total_pairs / g_s
        if args.gather == "auto":
            dev_s_for_line = dev_s            # headline stays the plain step; the gathered step is reported beside it
        else:
            value = gather["value"]
            dev_s_for_line = g_s
        gdist.close_peer_fields()
    else:
        dev_s_for_line = dev_s

    # ---- end to end through the public numpy API (host in, host out) ----
    out = None
    for _ in range(max(args.warmup, 3)):     # also warms the pinned-host allocator
        del out
        out = step_host()
    d2h = out.nbytes
    del out
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = step_host()
        del out
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    e2e_value = args.steps * total_pairs / e2e_s

    if rank == 0:
        # dominant kernel: per-launch duration from events recorded around it on its stream
        units = cfg["cov"].shape[0] if batched else (len(cfg["axes"][0]) if structured else cfg["pos"].shape[1])
        share = (hi - lo) / units     # rank 0's part of the sharded workload
        pairs_per_launch = total_pairs * share * args.steps / max(kern_n, 1)
        flop_per_pair = 2.0 * cfg["work_per_pair_dfma"]
        kern_s = kern_ms * 1e-3 / max(kern_n, 1)
        achieved = pairs_per_launch * flop_per_pair / kern_s / 1e12 if kern_n else None
        peak_tflops = 2.0 * peak_dfma / 1e12
        traffic, ncu_clock = None, None
        prof = os.path.join(REPO, "profiles", "dominant_kernel_traffic.json")
        if os.path.exists(prof):
            try:
                pj = json.load(open(prof))
                traffic, ncu_clock = pj.get(cfg["tag"]), pj.get("sm__cycles_elapsed.avg.per_second")
            except Exception:
                traffic = None
        roofline = {
            # DMMA.8x8x4 issues on the tensor pipe (ncu: sm__pipe_tensor_cycles_active), the direct
            # kernel's DFMAs on the FP64 pipe; both have the same measured fp64 peak
            "bound": "tensor" if structured else "fp64",
            "bound_detail": ("fp64 tensor path (DMMA.8x8x4), FP64-pipe roofline of BASELINE.json" if structured
                             else "FP64 pipe (DFMA), FP64-pipe roofline of BASELINE.json"),
            "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s",
            "frac": (achieved / peak_tflops) if achieved else None, "traffic": traffic,
            "kernel": "sk_contract_kernel (DMMA.8x8x4, stream-K, one launch per step)" if structured
                      else "direct_kernel",
            "kernel_ms_per_launch": 1e3 * kern_s, "launches_timed": kern_n,
            # (plan mode: the launches of all devices are summed, they run side by side)
            "kernel_share_of_step": (kern_ms * 1e-3) / dev_s / (1 if plan is None else len(plan)) if dev_s else None,
            "algorithmic_flop_per_pair": flop_per_pair,
            "peak_source": "DFMA microbenchmark run in this process (gsb_measure_fp64_peak); "
                           "MEASURED_PEAKS.json has no fp64 entry",
            "frac_of_nominal_37.2_TFLOPs": (achieved / (2 * NOMINAL_DFMA_PER_S / 1e12)) if achieved else None,
            # both fp64 paths, measured in this process: register-resident DFMA chains and DMMA.8x8x4 tiles, each at
            # 64 warps per SM (8 CTAs x 8 warps).  tools/microbench/fp64_pipe.cu tops out at 91.8 % with DFMA chains
            # because it runs at the contraction's occupancy (1-3 warps per sub-partition): a DFMA can only issue
            # every other cycle per warp, so few warps cannot fill the pipe; 16 warps per sub-partition can (99 %).
            "peaks_tflops": {"dfma": 2.0 * peak_dfma / 1e12, "dmma_8x8x4": 2.0 * peak_dmma / 1e12,
                             "nominal_148x64x1.965GHz": 2 * NOMINAL_DFMA_PER_S / 1e12},
            "ncu_sm_clock_hz": ncu_clock,
        }
        if "contract_per_pair_dfma" in cfg and achieved:
            # the same throughput against SURVEY 8(d)'s instruction BUDGET (a note, not the headline: the kernel
            # does the budgeted work in fewer instructions, it does not skip any)
            roofline["frac_vs_survey_contract_budget"] = roofline["frac"] * cfg["contract_per_pair_dfma"] / cfg["work_per_pair_dfma"]
        cpu = cpu_baseline(cfg) if (world == 1 and not args.no_cpu) else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world if plan is None else len(plan),
            "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dev_s_for_line / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(cfg, args),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * e2e_s / args.steps,
                    "d2h_gbs_per_gpu": d2h * args.steps / e2e_s / 1e9 / (1 if plan is None else len(plan)),
                    "d2h_gbs_all_gpus": d2h * args.steps / e2e_s / 1e9 * (world if plan is None else 1),
                    "api": "gstools_b200.summate_structured(numpy...) -> numpy (pinned)"
                           if structured else "gstools_b200.summate(numpy...) -> numpy"},
            "gpu_launches": int(launches), "roofline": roofline, "clocks": clocks.summary(),
            "wall_s_timed_region": t_wall,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if gather is not None:
            line["gather"] = gather
        if plan is not None:
            line["config"]["sharding"] = (f"ONE process, gsb_plan over {len(plan)} GPUs: device route stores every GPU's "
                                          "share into the result tensor on GPU 0 through peer memory, host route copies "
                                          "every share into its slice of one pinned array")
            line["config"]["peer_access"] = plan.peer_access
        print(json.dumps(line), flush=True)
    if plan is not None:
        plan.close()
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------
# widened row f1: the kriging evaluation of BASELINE.json configs[4] (--workload krige)
# ----------------------------------------------------------------------------------------------
KRIGE_METRIC = "Krige evaluation points/s (fp64, 1000 conditioning points, field + variance)"


def make_krige_workload(edge=128, n_cond=1000):
    """Ordinary kriging of config 5: 1000 synthetic conditioning points, Exponential(dim=3,
    var=1, len_scale=10), structured edge^3 mesh.  The kriging matrix is assembled and pseudo-inverted on
    the host exactly as Krige._get_krige_mat does (src/gstools/krige/base.py:330-357)."""
    rs = np.random.RandomState(20170519)
    cond_pos = rs.uniform(0, edge - 1, (3, n_cond))
    cond_val = rs.normal(size=n_cond)
    diff = cond_pos[:, :, None] - cond_pos[:, None, :]
    size = n_cond + 1
    mat = np.ones((size, size))
    mat[:n_cond, :n_cond] = np.exp(-np.sqrt((diff * diff).sum(axis=0)) / 10.0)
    mat[n_cond, n_cond] = 0.0
    return dict(spec=dict(kind="Exponential", var=1.0, len_rescaled=10.0), mat=np.linalg.pinv(mat),
                cond=np.concatenate([cond_val, [0.0]]), cond_pos=cond_pos, axes=[np.arange(float(edge))] * 3,
                size=size, n=edge ** 3, name=f"C5 kriging step: Ordinary kriging, {n_cond} points, "
                                              f"Exponential 3D, {edge}^3 structured mesh")


def krige_cpu(w, n_pts, offset=0):
    """The reference's chunk step on the host: right-hand sides (numpy) + native loop nest (oracle port)."""
    import oracle

    idx = (np.arange(n_pts) + offset) % w["n"]
    pos = bc.grid_points(w["axes"], None, idx)
    t0 = time.perf_counter()
    oracle.krige_evaluate(w["spec"], w["mat"], w["cond"], w["cond_pos"], pos, num_threads=host_threads())
    return time.perf_counter() - t0


def krige_config(w, world):
    return {"workload": w["name"], "baseline_config": "C5 (kriging step, SURVEY.md 8f row f1)",
            "path": "krige_evaluate (device-generated right-hand sides, triangular DMMA contraction)",
            "krige_size": int(w["size"]), "dim": 3, "sharding": f"axis-0 slabs over {world} GPU(s), no collective",
            "l2": "256 MiB flush buffer written between timed steps, outside the event brackets"}


def run_krige_reference_arm(args, w):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import oracle

    oracle.build()
    t = krige_cpu(w, 512)
    budget = min(20.0, 150.0 / max(1, args.steps + args.warmup))
    m = int(max(512, min(65536, 512 * budget / max(t, 1e-9))))
    for k in range(args.warmup):
        krige_cpu(w, m, offset=k * m)
    total = sum(krige_cpu(w, m, offset=(args.warmup + k) * m) for k in range(args.steps))
    value = args.steps * m / total
    sample = (f"each step = {m} contiguous mesh nodes on the host CPU: numpy right-hand sides "
              f"(Krige._get_krige_vecs) + C/OpenMP port of the native loop nest; linear in the number of points")
    print(json.dumps({
        "impl": "reference", "metric": KRIGE_METRIC, "value": value, "unit": "points/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": krige_config(w, args.gpus),
        "cpu_baseline": {"value": value, "unit": "points/s", "cores": host_threads(), "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}), flush=True)


def run_krige_arm(args, w):
    import torch
    import torch.distributed as dist

    import gstools_b200 as gsb
    from gstools_b200.dist import shard_range

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 backend has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    gsb.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    lo, hi = shard_range(len(w["axes"][0]), rank, world)
    h_axes = [np.ascontiguousarray(w["axes"][0][lo:hi])] + [np.ascontiguousarray(a) for a in w["axes"][1:]]
    n_local = int(np.prod([len(a) for a in h_axes]))
    t = lambda a: torch.tensor(np.ascontiguousarray(a), device=dev)
    d_mat, d_cond, d_cpos, d_axes = t(w["mat"]), t(w["cond"]), t(w["cond_pos"]), [t(a) for a in h_axes]

    def step_device():
        return gsb.krige_evaluate(w["spec"], d_mat, d_cond, d_cpos, axes=d_axes)

    def step_host():
        return gsb.krige_evaluate(w["spec"], w["mat"], w["cond"], w["cond_pos"], axes=h_axes)

    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    peak_dfma = gsb.measure_fp64_peak(local, 0, 0.4)
    for _ in range(args.warmup):
        step_device()
    barrier()
    gsb.set_option("time_kernels", 1)
    gsb.kernel_times()
    launches0 = gsb.get_counter("launches")
    events = []
    with ClockSampler(local) as clocks:
        t_wall0 = time.perf_counter()
        for _ in range(args.steps):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = step_device()
            e1.record()
            events.append((e0, e1))
            del out
        barrier()
        t_wall = time.perf_counter() - t_wall0
    launches = gsb.get_counter("launches") - launches0
    kern_ms, kern_n = gsb.kernel_times()
    gsb.set_option("time_kernels", 0)
    dev_s = sum(a.elapsed_time(b) for a, b in events) * 1e-3
    red = torch.tensor([dev_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
    dev_s = float(red.item())
    value = args.steps * w["n"] / dev_s
    for _ in range(3):
        step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = step_host()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    red = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
    e2e_s = float(red.item())
    if rank == 0:
        K = w["size"]
        fma_per_point = 0.5 * K * (K + 1) + K              # triangular quadratic form + the field row
        kern_s = kern_ms * 1e-3 / max(kern_n, 1)
        pts_per_launch = n_local * args.steps / max(kern_n, 1)
        achieved = pts_per_launch * 2.0 * fma_per_point / kern_s / 1e12 if kern_n else None
        peak_tflops = 2.0 * peak_dfma / 1e12
        traffic = None
        try:
            traffic = json.load(open(os.path.join(REPO, "profiles", "dominant_kernel_traffic.json"))).get("krige")
        except Exception:
            pass
        roofline = {
            "bound": "tensor", "bound_detail": "fp64 tensor path (DMMA.8x8x4), FP64-pipe roofline",
            "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s",
            "frac": (achieved / peak_tflops) if achieved else None, "traffic": traffic,
            "kernel": "krige_kernel<TILED> (DMMA.8x8x4)", "kernel_ms_per_launch": 1e3 * kern_s,
            "launches_timed": kern_n, "kernel_share_of_step": (kern_ms * 1e-3) / dev_s * (1 if world == 1 else 1),
            "algorithmic_flop_per_point": 2.0 * fma_per_point,
            "reference_loop_flop_per_point": 2.0 * K * K,
            "peak_source": "DFMA microbenchmark run in this process (gsb_measure_fp64_peak); "
                           "MEASURED_PEAKS.json has no fp64 entry"}
        line = {
            "metric": KRIGE_METRIC, "value": value, "unit": "points/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dev_s / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": krige_config(w, world),
            "e2e": {"value": args.steps * w["n"] / e2e_s, "unit": "points/s",
                    "h2d_bytes_per_step": int(w["mat"].nbytes + w["cond"].nbytes + w["cond_pos"].nbytes
                                              + sum(a.nbytes for a in h_axes)),
                    "d2h_bytes_per_step": int(2 * out[0].nbytes), "ms_per_step": 1e3 * e2e_s / args.steps,
                    "api": "gstools_b200.krige_evaluate(numpy...) -> (field, error) numpy"},
            "gpu_launches": int(launches), "roofline": roofline, "clocks": clocks.summary(),
            "wall_s_timed_region": t_wall}
        if world == 1:
            import oracle

            oracle.build()
            tp = krige_cpu(w, 256)
            m = int(max(256, min(16384, 256 * 10.0 / max(tp, 1e-9))))
            tc = krige_cpu(w, m, offset=m)
            line["cpu_baseline"] = {"value": m / tc, "unit": "points/s", "cores": host_threads(), "kind": "port",
                                    "sample": f"{m} contiguous mesh nodes ({tc:.1f} s): numpy right-hand sides + "
                                              f"C/OpenMP port of the native loop nest"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------
# BASELINE.json configs[4] through the drop-in path: gs.CondSRF realisations with the plugin (--workload c5cond)
# ----------------------------------------------------------------------------------------------
def run_c5cond_arm(args, edge=128, n_cond=1000):
    """A "step" is ONE conditioned realisation through the unmodified reference's ``gs.CondSRF.__call__`` with
    ``gstools_b200.enable()`` -- seed -> mode set (host) -> summation with the kriging results fused into the
    kernel's stores (device) -> conditioned field as a host array.  Host in, host out: ``value`` and ``e2e``
    are the same wall-clock measurement here (there is no device-resident variant of this call).  The
    kriging system (ordinary kriging, 1000 points, Exponential 3D) is evaluated once, in the first warm-up
    step, and reported separately."""
    import torch

    sys.path.insert(0, os.path.join(REPO, "tests"))
    import refharness

    if not refharness.have_reference():
        raise SystemExit("bench.py --workload c5cond needs the reference gstools (baseline/_ref; run "
                         "__graft_entry__.build() in the build container)")
    gs = refharness.import_gstools()
    import gstools_b200 as gsb

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 backend has no CPU fallback)")
    # N > 1 (torchrun): the ensemble's seeds are sharded over the ranks, one process per GPU (BASELINE.json
    # configs[4]: "seeds sharded over 8 GPUs"); every rank sets up the (replicated) kriging system itself and runs
    # `steps` realisations of ITS seeds -- nothing is exchanged.  Weak scaling: the work per GPU is fixed.
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    gsb.set_device(local)
    gsb.enable()
    rs = np.random.RandomState(20170519)
    cond_pos = rs.uniform(0, edge - 1, (3, n_cond))
    cond_val = rs.normal(size=n_cond)
    model = gs.Exponential(dim=3, var=1, len_scale=10)
    t0 = time.perf_counter()
    krige = gs.krige.Ordinary(model, cond_pos, cond_val)
    t_setup = time.perf_counter() - t0
    crf = gs.CondSRF(krige)
    crf.set_pos([np.arange(float(edge))] * 3, "structured")
    master = gs.random.MasterRNG(20170519)
    all_seeds = [master() for _ in range((max(args.warmup, 3) + args.steps + 8) * world)]
    my_seeds = iter(all_seeds[rank::world])                  # this rank's share of the ensemble's seeds

    def seeds():
        return next(my_seeds)

    n_modes = 1000
    t0 = time.perf_counter()
    crf(seed=seeds(), store=["fld", False, False])           # evaluates the kriging system
    t_first = time.perf_counter() - t0
    for _ in range(max(args.warmup, 3) - 1):
        crf(seed=seeds(), store=["fld", False, False])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches0 = gsb.get_counter("launches")
    k0 = gsb.get_counter("krige_calls")
    t_modes = 0.0
    with ClockSampler(local) as clocks:
        t0 = time.perf_counter()
        for _ in range(args.steps):
            f = crf(seed=seeds(), store=["fld", False, False])
        torch.cuda.synchronize()
        total = time.perf_counter() - t0
    if world > 1:
        tt = torch.tensor([total], dtype=torch.float64, device=torch.device("cuda", local))
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total = float(tt.item())
    launches = gsb.get_counter("launches") - launches0
    assert gsb.get_counter("krige_calls") == k0, "the kriging system must not be re-evaluated per realisation"
    # share of the host-side mode sampling (RandMeth.reset_seed with the native radius sampler)
    t0 = time.perf_counter()
    for _ in range(8):
        crf.generator.update(model, seeds())
    t_modes = (time.perf_counter() - t0) / 8
    pairs = edge ** 3 * n_modes
    value = world * args.steps * pairs / total
    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        gsb.disable()
        return
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"C5 as configured: gs.CondSRF realisations, Ordinary kriging on {n_cond} points, "
                               f"Exponential 3D, {edge}^3 structured mesh, mode_no={n_modes}, plugin enabled",
                   "baseline_config": "C5", "path": "CondSRF.__call__ -> fused per-point epilogue (scaled contraction)",
                   "mode_no": n_modes, "dim": 3,
                   "step": "one conditioned realisation per GPU, seed in -> host field out",
                   "sharding": f"seeds of the ensemble dealt out to {world} process(es), one per GPU, no collective",
                   "l2": "every realisation writes a fresh 16.8 MB field and reads 33.6 MB of kriging results; "
                         "inputs are re-sampled per step"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": int(8 * (5 * n_modes + 3 * edge)),
                "d2h_bytes_per_step": int(3 * 8 * edge ** 3), "ms_per_step": 1e3 * total / args.steps,
                "api": "gs.CondSRF(krige)(seed=s, store=[name, False, False]) with gstools_b200.enable()"},
        "gpu_launches": int(launches), "clocks": clocks.summary(),
        "breakdown_ms": {"kriging_setup_pinv_host": 1e3 * t_setup, "first_call_incl_kriging_evaluation": 1e3 * t_first,
                         "mode_sampling_host_per_realisation": 1e3 * t_modes,
                         "ensemble_of_256_extrapolated_s": t_first + (256 / world - 1) * total / args.steps},
        "roofline": None, "wall_s_timed_region": total,
    }
    print(json.dumps(line), flush=True)
    gsb.disable()


# ----------------------------------------------------------------------------------------------
# BASELINE.json configs[4] as ONE call: gstools_b200.ensemble(cond_srf, 256 seeds) (--workload c5ens)
# ----------------------------------------------------------------------------------------------
def run_c5ens_arm(args, edge=128, n_cond=1000, n_seeds=256):
    """A "step" is the WHOLE ensemble: 256 seeds from MasterRNG(20170519) -> 256 conditioned 128^3 fields as one host
    array, through ``gstools_b200.ensemble(gs.CondSRF(krige), seeds)``: mode sets of all seeds drawn natively on the
    host cores, one batched summation with the kriging results fused into the stores (``--plan G``: seeds dealt out to
    G GPUs from this one process), 4.3 GB copied to the host.  Host in, host out: value == e2e."""
    import torch

    sys.path.insert(0, os.path.join(REPO, "tests"))
    import refharness

    if not refharness.have_reference():
        raise SystemExit("bench.py --workload c5ens needs the reference gstools (baseline/_ref)")
    if int(os.environ.get("RANK", "0")) != 0:
        return
    gs = refharness.import_gstools()
    import gstools_b200 as gsb

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 backend has no CPU fallback)")
    gsb.set_device(0)
    gsb.enable()
    n_gpus = 1
    if args.plan > 1:
        gsb.use_devices(list(range(args.plan)), min_pairs=0)
        n_gpus = args.plan
    rs = np.random.RandomState(20170519)
    cond_pos = rs.uniform(0, edge - 1, (3, n_cond))
    cond_val = rs.normal(size=n_cond)
    model = gs.Exponential(dim=3, var=1, len_scale=10)
    krige = gs.krige.Ordinary(model, cond_pos, cond_val)
    crf = gs.CondSRF(krige)
    axes = [np.arange(float(edge))] * 3
    crf.set_pos(axes, "structured")
    master = gs.random.MasterRNG(20170519)
    seeds = [master() for _ in range(n_seeds)]
    n_modes = 1000
    t0 = time.perf_counter()
    out = gsb.ensemble(crf, seeds[:8])            # evaluates the kriging system, warms pools
    t_first = time.perf_counter() - t0
    del out
    for _ in range(max(1, args.warmup - 2)):
        out = gsb.ensemble(crf, seeds)
        del out
    t0 = time.perf_counter()
    gsb.sample_modes_batch("Exponential", 3, model.len_rescaled, 0.0, seeds, n_modes)
    t_modes = time.perf_counter() - t0
    launches0 = gsb.get_counter("launches")
    steps = max(1, min(args.steps, 5))
    with ClockSampler(0) as clocks:
        t0 = time.perf_counter()
        for _ in range(steps):
            out = gsb.ensemble(crf, seeds)
            nbytes = out.nbytes
            del out
        torch.cuda.synchronize()
        total = time.perf_counter() - t0
    launches = gsb.get_counter("launches") - launches0
    pairs = n_seeds * edge ** 3 * n_modes
    value = steps * pairs / total
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": steps, "warmup": max(1, args.warmup - 2),
        "ms_per_step": 1e3 * total / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"C5 as ONE call: gstools_b200.ensemble(gs.CondSRF(Ordinary kriging on {n_cond} points, "
                               f"Exponential 3D), {n_seeds} seeds from MasterRNG(20170519)) on the {edge}^3 mesh, mode_no={n_modes}",
                   "baseline_config": "C5", "path": "native batch sampler -> batched stream-K contraction with the per-point epilogue",
                   "mode_no": n_modes, "dim": 3, "step": f"the whole ensemble of {n_seeds} realisations, seeds in -> host array out",
                   "sharding": ("one GPU" if n_gpus == 1 else f"ONE process, seeds dealt out to {n_gpus} GPUs (gsb_plan)"),
                   "l2": "4.3 GB of output per step"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": int(8 * 5 * n_modes * n_seeds),
                "d2h_bytes_per_step": int(nbytes), "ms_per_step": 1e3 * total / steps,
                "api": "gstools_b200.ensemble(gs.CondSRF(krige), seeds) -> numpy (pinned)"},
        "gpu_launches": int(launches), "clocks": clocks.summary(),
        "breakdown_ms": {"first_call_8_seeds_incl_kriging_evaluation": 1e3 * t_first,
                         "mode_sets_of_all_seeds_native_batch": 1e3 * t_modes, "host_threads": host_threads()},
        "roofline": None, "wall_s_timed_region": total,
    }
    print(json.dumps(line), flush=True)
    gsb.use_devices(None)
    gsb.disable()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c1", "c2", "c3", "c4", "c5", "c5cond", "c5ens", "krige"])
    ap.add_argument("--gather", default="auto", choices=["auto", "none", "nccl", "p2p"],
                    help="N > 1, structured workloads: deliver the field as one device array on rank 0.  nccl / p2p: "
                         "`value` is compute + gather; auto (default): `value` is the plain step and the p2p-gathered "
                         "step is measured and reported beside it under \"gather\"; none: skip")
    ap.add_argument("--pieces", type=int, default=1, help="row pieces per rank of --gather nccl")
    ap.add_argument("--reserve-sms", type=int, default=0, help="SMs left to NCCL while pieces travel (--gather nccl)")
    ap.add_argument("--plan", type=int, default=0, help="single-process mode: drive this many GPUs through gsb_plan")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (builder's table runs)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.workload == "krige":
        w = make_krige_workload()
        return run_krige_reference_arm(args, w) if args.impl == "reference" else run_krige_arm(args, w)
    if args.workload == "c5cond":
        if args.impl == "reference":
            return run_reference_arm(args, make_workload("c5"))
        return run_c5cond_arm(args)
    if args.workload == "c5ens":
        if args.impl == "reference":
            return run_reference_arm(args, make_workload("c5"))
        return run_c5ens_arm(args)
    cfg = make_workload(args.workload)
    if args.impl == "reference":
        run_reference_arm(args, cfg)
    else:
        run_b200_arm(args, cfg)


if __name__ == "__main__":
    main()
