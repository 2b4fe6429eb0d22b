"""How fast can the GPUs of this box deliver results into host memory AT THE SAME TIME?

The end-to-end numbers of a multi-GPU run are bound by this (8 ranks delivered 1.07 GB at ~98 GB/s
aggregate although one GPU alone reaches ~53 GB/s).  For k = 1, 2, 4, ... GPUs: every GPU copies a 512 MiB device
buffer into (a) its own pinned buffer, (b) its slice of ONE pinned buffer, (c) its slice of one ordinary numpy array
registered with cudaHostRegister; all copies are started together, the wall time of the slowest counts.
    python tools/d2h_probe.py            -> one JSON line per (k, variant)
"""
import json
import sys
import time

import numpy as np
import torch

MB = 512


def run(k, variant, reps=5):
    n = MB * (1 << 20) // 8
    devs = [torch.device("cuda", i) for i in range(k)]
    src = [torch.ones(n, dtype=torch.float64, device=d) for d in devs]
    streams = [torch.cuda.Stream(device=d) for d in devs]
    if variant == "own_pinned":
        dst = [torch.empty(n, dtype=torch.float64).pin_memory() for _ in devs]
    elif variant == "one_pinned":
        big = torch.empty(n * k, dtype=torch.float64).pin_memory()
        dst = [big[i * n:(i + 1) * n] for i in range(k)]
    else:
        arr = np.empty(n * k, dtype=np.float64)
        arr[:] = 0.0
        big = torch.from_numpy(arr)
        rc = torch.cuda.cudart().cudaHostRegister(big.data_ptr(), big.numel() * 8, 0)
        assert int(rc) == 0, rc
        dst = [big[i * n:(i + 1) * n] for i in range(k)]
    best = 1e9
    for _ in range(reps):
        for d in devs:
            torch.cuda.synchronize(d)
        t0 = time.perf_counter()
        for i in range(k):
            with torch.cuda.stream(streams[i]):
                dst[i].copy_(src[i], non_blocking=True)
        for d in devs:
            torch.cuda.synchronize(d)
        best = min(best, time.perf_counter() - t0)
    if variant == "registered":
        torch.cuda.cudart().cudaHostUnregister(big.data_ptr())
    gb = MB * (1 << 20) / 1e9
    return {"gpus": k, "variant": variant, "gbs_aggregate": k * gb / best, "gbs_per_gpu": gb / best,
            "ms": best * 1e3}


def run_write_combined(k, reps=5):
    """Destination allocated with cudaHostAllocWriteCombined: the CPU caches do not have to be snooped for the
    incoming writes (reading such memory from the CPU is slow -- a probe of what limits the host, not a product option)."""
    import ctypes

    rt = None
    for name in ("libcudart.so.12", "libcudart.so"):
        try:
            rt = ctypes.CDLL(name)
            break
        except OSError:
            continue
    if rt is None:
        return None
    n_bytes = MB * (1 << 20)
    devs = [torch.device("cuda", i) for i in range(k)]
    src = [torch.ones(n_bytes // 8, dtype=torch.float64, device=d) for d in devs]
    streams = [torch.cuda.Stream(device=d) for d in devs]
    dst = []
    for _ in devs:
        p = ctypes.c_void_p()
        if rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(n_bytes), ctypes.c_uint(0x04 | 0x01)) != 0:   # WC | portable
            return None
        dst.append(p)
    best = 1e9
    for _ in range(reps):
        for d in devs:
            torch.cuda.synchronize(d)
        t0 = time.perf_counter()
        for i in range(k):
            torch.cuda.set_device(devs[i])
            rt.cudaMemcpyAsync(dst[i], ctypes.c_void_p(src[i].data_ptr()), ctypes.c_size_t(n_bytes), ctypes.c_int(2),
                               ctypes.c_void_p(streams[i].cuda_stream))
        for d in devs:
            torch.cuda.synchronize(d)
        best = min(best, time.perf_counter() - t0)
    for p in dst:
        rt.cudaFreeHost(p)
    gb = n_bytes / 1e9
    return {"gpus": k, "variant": "write_combined_pinned", "gbs_aggregate": k * gb / best, "gbs_per_gpu": gb / best,
            "ms": best * 1e3}


def numa_nodes():
    import glob

    return sorted(int(p.rsplit("node", 1)[1]) for p in glob.glob("/sys/devices/system/node/node[0-9]*"))


def interleave_policy(nodes):
    """set_mempolicy(MPOL_INTERLEAVE) for this thread: pages of later (pinned) allocations alternate over `nodes`."""
    import ctypes

    mask = ctypes.c_ulong(sum(1 << n for n in nodes))
    libc = ctypes.CDLL(None, use_errno=True)
    rc = libc.syscall(238, 3, ctypes.byref(mask), ctypes.c_ulong(65))       # SYS_set_mempolicy, MPOL_INTERLEAVE
    return rc == 0


def main():
    nodes = numa_nodes()
    print(json.dumps({"numa_nodes_visible": nodes}), flush=True)
    total = torch.cuda.device_count()
    ks = [k for k in (1, 2, 4, 8) if k <= total]
    for k in ks:
        for variant in ("own_pinned", "one_pinned", "registered"):
            print(json.dumps(run(k, variant)), flush=True)
    for k in ks:
        try:
            r = run_write_combined(k)
        except Exception as exc:  # noqa: BLE001
            r = {"gpus": k, "variant": "write_combined_pinned", "error": str(exc)}
        if r is not None:
            print(json.dumps(r), flush=True)
    if len(nodes) > 1 and interleave_policy(nodes):
        for k in ks:
            r = run(k, "own_pinned")
            r["variant"] = "own_pinned_numa_interleaved"
            print(json.dumps(r), flush=True)
    # H2D for comparison
    n = MB * (1 << 20) // 8
    for k in ks:
        devs = [torch.device("cuda", i) for i in range(k)]
        dst = [torch.empty(n, dtype=torch.float64, device=d) for d in devs]
        src = [torch.ones(n, dtype=torch.float64).pin_memory() for _ in devs]
        streams = [torch.cuda.Stream(device=d) for d in devs]
        best = 1e9
        for _ in range(5):
            for d in devs:
                torch.cuda.synchronize(d)
            t0 = time.perf_counter()
            for i in range(k):
                with torch.cuda.stream(streams[i]):
                    dst[i].copy_(src[i], non_blocking=True)
            for d in devs:
                torch.cuda.synchronize(d)
            best = min(best, time.perf_counter() - t0)
        gb = MB * (1 << 20) / 1e9
        print(json.dumps({"gpus": k, "variant": "h2d_own_pinned", "gbs_aggregate": k * gb / best,
                          "gbs_per_gpu": gb / best, "ms": best * 1e3}), flush=True)


if __name__ == "__main__":
    sys.exit(main())
