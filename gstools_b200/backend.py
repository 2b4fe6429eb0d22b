"""Host-side mirror of the reference's native summator interface.

Same names, argument order and meaning as the functions gstools imports from
gstools-cython / gstools_core (reference: src/gstools/field/generator.py:22-34)
and calls at generator.py:48 and :64::

    summate(cov_samples, z_1, z_2, pos, num_threads=None)         -> (n,)   float64
    summate_incompr(cov_samples, z_1, z_2, pos, num_threads=None) -> (d, n) float64

``num_threads`` is accepted for signature compatibility and ignored.

Inputs may be numpy arrays (host path: the C ABI stages the copies, the result is a
numpy array in pinned memory) or CUDA ``torch`` tensors (device path: zero copy, work
is enqueued on the current torch stream, the result is a CUDA tensor).

Everything runs through ``libgsb200.so``; there is no CPU fallback.
"""

from __future__ import annotations

import ctypes
import os
import weakref

import numpy as np

from . import _lib

__all__ = [
    "summate",
    "summate_incompr",
    "summate_structured",
    "summate_incompr_structured",
    "summate_fourier",
    "summate_fourier_structured",
    "calc_field_krige_and_variance",
    "calc_field_krige",
    "krige_evaluate",
    "cov_model_spec",
    "sample_radii_mcmc",
    "sample_modes_batch",
    "scale_shift_",
    "make_epilogue",
    "make_point_epilogue",
    "cond_scaling",
    "to_device",
    "to_host",
    "to_host_async",
    "set_device",
    "get_device",
    "Plan",
    "use_devices",
    "current_plan",
    "structured_slab_into",
]

_DEVICE = None
_PIN_THRESHOLD = 1 << 16  # outputs of at least this many doubles are allocated pinned


def set_device(index: int):
    """Select the CUDA device used for numpy (host-buffer) calls."""
    global _DEVICE
    _DEVICE = int(index)


def get_device() -> int:
    if _DEVICE is not None:
        return _DEVICE
    env = os.environ.get("GSB200_DEVICE")
    if env is not None:
        return int(env)
    if "LOCAL_RANK" in os.environ:  # one process per GPU under torchrun
        # launchers that narrow CUDA_VISIBLE_DEVICES per rank (SLURM, one visible GPU per process)
        # leave fewer visible devices than local ranks: wrap instead of failing with "invalid device"
        rank = int(os.environ["LOCAL_RANK"])
        try:
            count = _lib.device_count()
        except Exception:
            count = 0
        return rank % count if count > 0 else rank
    return 0


# ----------------------------------------------------------------------------------------
# helpers
# ----------------------------------------------------------------------------------------
def _is_cuda_tensor(x) -> bool:
    return type(x).__module__.startswith("torch") and getattr(x, "is_cuda", False)


def _torch():
    import torch

    return torch


_PINNED = {"bytes": 0, "limit": None}


def _pinned_limit():
    """Cap on the pinned host memory held by arrays this module handed out (they stay pinned for as long
    as the caller keeps them): ``GSB200_PINNED_LIMIT_MB``, default a quarter of the physical memory."""
    if _PINNED["limit"] is None:
        env = os.environ.get("GSB200_PINNED_LIMIT_MB")
        if env is not None:
            _PINNED["limit"] = int(float(env) * (1 << 20))
        else:
            try:
                total = os.sysconf("SC_PAGE_SIZE") * os.sysconf("SC_PHYS_PAGES")
            except (ValueError, OSError):
                total = 64 << 30
            _PINNED["limit"] = total // 4
    return _PINNED["limit"]


def _release_pinned(nbytes):
    _PINNED["bytes"] -= nbytes


def _empty_host(shape):
    """Uninitialised float64 host array; pinned when large so D2H runs at full PCIe speed."""
    n = int(np.prod(shape)) if len(shape) else 1
    nbytes = 8 * n
    if n >= _PIN_THRESHOLD and _PINNED["bytes"] + nbytes <= _pinned_limit():
        try:
            torch = _torch()
            if torch.cuda.is_available():
                tensor = torch.empty(tuple(shape), dtype=torch.float64, pin_memory=True)
                arr = tensor.numpy()
                _PINNED["bytes"] += nbytes
                weakref.finalize(arr, _release_pinned, nbytes)   # views keep `arr` alive through .base
                return arr
        except Exception:  # pinned allocation is an optimisation only
            pass
    return np.empty(shape, dtype=np.float64)


def to_device(a, device=None):
    """Contiguous float64 CUDA tensor holding a copy of the host array ``a`` (None passes through)."""
    if a is None:
        return None
    torch = _torch()
    dev = torch.device("cuda", get_device() if device is None else int(device))
    return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64), device=dev)


def to_host(t):
    """Fresh host array (pinned when large, so the copy runs at PCIe speed) with the contents of CUDA tensor ``t``."""
    torch = _torch()
    out = _empty_host(tuple(t.shape))
    if out.size:
        torch.from_numpy(out).copy_(t)
    return out


_COPY_STREAMS = {}


def to_host_async(t):
    """Start copying CUDA tensor ``t`` into a fresh (pinned) host array on a side stream and return
    ``(array, event)``: the array holds the data once ``event.synchronize()`` has returned.  Lets a D2H copy travel
    while the host does something else (the plugin overlaps the krige-side results of a conditioned realisation with
    the sampling of its mode set)."""
    torch = _torch()
    out = _empty_host(tuple(t.shape))
    key = t.device.index
    side = _COPY_STREAMS.get(key)
    if side is None:
        side = _COPY_STREAMS[key] = torch.cuda.Stream(device=t.device)
    side.wait_stream(torch.cuda.current_stream(t.device))
    event = torch.cuda.Event()
    with torch.cuda.stream(side):
        if out.size:
            torch.from_numpy(out).copy_(t, non_blocking=True)
        event.record(side)
    return out, event


def _as_f64(a, name):
    try:
        arr = np.asarray(a, dtype=np.float64)
    except (TypeError, ValueError) as exc:
        raise TypeError(f"{name}: cannot be interpreted as a float64 array") from exc
    return arr


def _rows_contiguous(a):
    """(array, leading dimension in elements) with unit inner stride; copies only when needed."""
    n = a.shape[1]
    if a.flags.c_contiguous:
        return a, max(n, 1)
    it = a.itemsize
    if n > 0 and a.strides[1] == it and a.strides[0] > 0 and a.strides[0] % it == 0 \
            and a.strides[0] // it >= n:
        return a, a.strides[0] // it  # row-strided view (e.g. a point range of a bigger array)
    return np.ascontiguousarray(a), max(n, 1)


def _check_modes(cov, z1, z2):
    if cov.ndim != 2:
        raise ValueError("cov_samples must have shape (dim, mode_no)")
    if z1.ndim != 1 or z2.ndim != 1 or z1.shape[0] != cov.shape[1] or z2.shape[0] != cov.shape[1]:
        raise ValueError("z_1 and z_2 must have shape (mode_no,) matching cov_samples")


def _ptr(a):
    return a.ctypes.data


def make_epilogue(scale, adds=()):
    """Build the fused caller epilogue ``v = scale*sum; v += adds[0]; v += adds[1]; ...``.

    Each entry of ``adds`` is a scalar or a per-component sequence (vector fields).  The kernels
    round every operation separately, in this order, like the numpy passes of the reference
    (generator.py:269-270, 561-567; normalizer/tools.py:99-103).  ``None`` means raw sums.
    """
    if scale is None:
        return None
    adds = list(adds)
    if len(adds) > _lib.EPI_MAX_ADD:
        raise ValueError(f"at most {_lib.EPI_MAX_ADD} additive terms can be fused")
    epi = _lib.Epilogue()
    epi.scale = float(scale)
    epi.n_add = len(adds)
    for k, a in enumerate(adds):
        vals = np.asarray(a, dtype=np.float64).reshape(-1)
        if vals.size not in (1, 2, 3):
            raise ValueError("an additive term is a scalar or one value per field component")
        for c in range(_lib.EPI_MAX_COMP):
            epi.add[k][c] = float(vals[c] if vals.size > 1 and c < vals.size else vals[0])
    return epi


class _PointEpi:
    """A ``gsb_point_epilogue`` together with the CUDA tensors it points into (kept alive with it)."""

    def __init__(self, gain, offset, adds):
        adds = [float(a) for a in adds]
        if len(adds) > _lib.EPI_MAX_ADD:
            raise ValueError(f"at most {_lib.EPI_MAX_ADD} additive terms can be fused")
        torch = _torch()
        self.tensors = []
        self.n = None
        self.device = None
        ptrs = []
        for name, t in (("gain", gain), ("offset", offset)):
            if t is None:
                ptrs.append(None)
                continue
            if not _is_cuda_tensor(t) or t.dtype != torch.float64 or not t.is_contiguous():
                raise TypeError(f"point epilogue: {name} must be a contiguous float64 CUDA tensor "
                                "(the device-resident result of a kriging evaluation)")
            if self.n is not None and t.numel() != self.n:
                raise ValueError("point epilogue: gain and offset must have the same number of points")
            if self.device is not None and t.device != self.device:
                raise ValueError("point epilogue: gain and offset must live on the same device")
            self.n, self.device = t.numel(), t.device
            self.tensors.append(t)
            ptrs.append(t.data_ptr())
        self.struct = _lib.PointEpilogue()
        self.struct.gain, self.struct.offset = ptrs
        self.struct.n_add = len(adds)
        for k, a in enumerate(adds):
            self.struct.add[k] = a

    def ref(self, n, device_index):
        if self.n is not None and self.n != n:
            raise ValueError(f"point epilogue: arrays hold {self.n} points, the field has {n}")
        if self.device is not None and self.device.index != device_index:
            raise ValueError("point epilogue: arrays live on another device than the call runs on")
        return ctypes.byref(self.struct)


def make_point_epilogue(gain=None, offset=None, adds=()):
    """Per-point step of a conditioned field (``gsb_point_epilogue``): after the terms of
    :func:`make_epilogue`, ``v = gain[i]*v; v = offset[i] + v; v += adds[0]; ...`` -- CondSRF's
    ``rawkrige + var_scale * rawfield + nugget`` (cond_srf.py:145-150) plus constant mean / trend, with
    separately rounded operations in that order.  ``gain`` / ``offset`` are contiguous float64 CUDA
    tensors with one entry per point of the field (device-resident kriging results), or None."""
    return _PointEpi(gain, offset, adds)


def cond_scaling(error, sill, var):
    """``(krige_var, gain)`` on the device from the kriging error sums (CUDA tensor in, CUDA tensors out):
    ``krige_var = max(sill - error, 0)`` (krige/base.py:296-298), ``gain = sqrt(krige_var / var)``
    (CondSRF.get_scaling without nugget, cond_srf.py:175-177), numpy's bits."""
    torch = _torch()
    if not _is_cuda_tensor(error) or error.dtype != torch.float64 or not error.is_contiguous():
        raise TypeError("cond_scaling works on contiguous float64 CUDA tensors")
    krige_var, gain = torch.empty_like(error), torch.empty_like(error)
    stream = torch.cuda.current_stream(error.device).cuda_stream
    rc = _lib.load().gsb_cond_scaling(error.data_ptr(), error.numel(), float(sill), float(var),
                                      krige_var.data_ptr(), gain.data_ptr(), error.device.index, stream)
    _lib.check(rc, "cond_scaling")
    return krige_var, gain


def _epi_ref(epilogue):
    if epilogue is None:
        return None
    if not isinstance(epilogue, _lib.Epilogue):
        epilogue = make_epilogue(*epilogue)
    return ctypes.byref(epilogue)


# ----------------------------------------------------------------------------------------
# multi-GPU plan: one process, several GPUs (gsb_plan_* of include/gsb200.h)
# ----------------------------------------------------------------------------------------
_PLAN = {"plan": None, "min_pairs": 4e9}


class _PlanPointEpi:
    """Per-device ``gsb_point_epilogue`` array of a :class:`Plan`: ``gain`` / ``offset`` replicated on every device."""

    def __init__(self, plan, gain, offset, adds):
        torch = _torch()
        self.plan = plan
        self.entries = []
        for d in plan.devices:
            dev = torch.device("cuda", d)
            rep = [None if t is None else torch.as_tensor(t, dtype=torch.float64).to(dev).contiguous().reshape(-1)
                   for t in (gain, offset)]
            self.entries.append(_PointEpi(rep[0], rep[1], adds))
        self.n = self.entries[0].n
        self.array = (_lib.PointEpilogue * len(self.entries))()
        for g, e in enumerate(self.entries):
            self.array[g] = e.struct

    def ref(self, n):
        if self.n is not None and self.n != n:
            raise ValueError(f"point epilogue: arrays hold {self.n} points, the field has {n}")
        return ctypes.cast(self.array, ctypes.c_void_p)


class Plan:
    """Several GPUs driven from ONE process (``gsb_plan_create``): a call is cut into independent shares --
    point ranges, slabs along axis 0 of a mesh, or batch entries (ensemble seeds) when there are at least as many
    as devices -- with no inter-GPU traffic during the sum.  Host arrays: every GPU copies its share straight
    into its slice of the one (pinned) result array.  CUDA tensors: the other GPUs store their share directly into
    the result tensor on the inputs' device through NVLink peer memory, from the kernels' epilogue.

    ``devices``: list of CUDA device indices, or None / "all" for every visible device.
    """

    def __init__(self, devices=None):
        lib = _lib.load()
        handle = ctypes.c_void_p()
        if devices is None or (isinstance(devices, str) and devices == "all"):
            arr, n = None, 0
        else:
            ids = [int(d) for d in devices]
            if not ids:
                raise ValueError("Plan: empty device list")
            arr, n = (ctypes.c_int * len(ids))(*ids), len(ids)
        _lib.check(lib.gsb_plan_create(arr, n, ctypes.byref(handle)), "plan_create")
        self._handle = handle
        cnt, peer = ctypes.c_int(0), ctypes.c_int(0)
        _lib.check(lib.gsb_plan_info(handle, ctypes.byref(cnt), None, None), "plan_info")
        devs = (ctypes.c_int * max(cnt.value, 1))()
        _lib.check(lib.gsb_plan_info(handle, ctypes.byref(cnt), devs, ctypes.byref(peer)), "plan_info")
        self.devices = [int(devs[i]) for i in range(cnt.value)]
        self.peer_access = bool(peer.value)
        self._finalizer = weakref.finalize(self, lib.gsb_plan_destroy, handle)

    def close(self):
        """Stop the plan's host threads (idempotent; also done when the object is collected)."""
        if _PLAN["plan"] is self:
            _PLAN["plan"] = None
        self._finalizer()
        self._handle = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __len__(self):
        return len(self.devices)

    @property
    def handle(self):
        """The ``gsb_plan *`` of this plan; using a closed plan is an error, not a crash."""
        if self._handle is None:
            raise ValueError("this Plan has been closed")
        return self._handle

    def share(self, n, part):
        """``[lo, hi)`` of the units device number ``part`` of the plan gets out of ``n``."""
        return _lib.plan_share(n, len(self.devices), part)

    def make_point_epilogue(self, gain=None, offset=None, adds=()):
        """:func:`make_point_epilogue` with ``gain`` / ``offset`` (host arrays or tensors covering all points of a
        field) replicated on every device of the plan."""
        return _PlanPointEpi(self, gain, offset, adds)

    def summate(self, cov_samples, z_1, z_2, pos, *, epilogue=None, point_epilogue=None):
        return _flat(cov_samples, z_1, z_2, pos, vec=False, epilogue=epilogue, point_epilogue=point_epilogue, plan=self)

    def summate_incompr(self, cov_samples, z_1, z_2, pos, *, epilogue=None):
        return _flat(cov_samples, z_1, z_2, pos, vec=True, epilogue=epilogue, plan=self)

    def summate_structured(self, cov_samples, z_1, z_2, axes, matrix=None, *, epilogue=None, point_epilogue=None):
        return _structured(cov_samples, z_1, z_2, axes, matrix, vec=False, epilogue=epilogue,
                           point_epilogue=point_epilogue, plan=self)

    def summate_incompr_structured(self, cov_samples, z_1, z_2, axes, matrix=None, *, epilogue=None):
        return _structured(cov_samples, z_1, z_2, axes, matrix, vec=True, epilogue=epilogue, plan=self)

    def krige_evaluate(self, model, krig_mat, cond, cond_pos, pos=None, axes=None, matrix=None, unbiased=True,
                       tail_rows=None, return_var=True):
        """:func:`krige_evaluate` with the evaluation points dealt out to the devices of the plan."""
        return krige_evaluate(model, krig_mat, cond, cond_pos, pos=pos, axes=axes, matrix=matrix, unbiased=unbiased,
                              tail_rows=tail_rows, return_var=return_var, plan=self)


def use_devices(devices="all", min_pairs=None):
    """Route the host-array entry points of this module (and with them ``gs.SRF`` / ``gs.CondSRF`` under
    :func:`gstools_b200.enable`) through a :class:`Plan` over ``devices`` whenever a call has at least
    ``min_pairs`` (point, mode) pairs (default 4e9: below that one GPU finishes before the fan-out pays).
    ``devices=None`` switches back to one device.  Returns the plan (or None)."""
    old = _PLAN["plan"]
    if min_pairs is not None:
        _PLAN["min_pairs"] = float(min_pairs)
    if devices is None:
        _PLAN["plan"] = None
        if old is not None:
            old.close()
        return None
    plan = Plan(devices)
    if len(plan.devices) < 2:
        plan.close()
        plan = None
    _PLAN["plan"] = plan
    if old is not None and old is not plan:
        old._finalizer()
    return plan


def current_plan():
    return _PLAN["plan"]


def _pick_plan(plan, pairs, point_epilogue, device_index=None):
    """The plan a call runs on: the explicit one, the one its point epilogue was built for, the process-wide one
    (big calls only), or None (one device)."""
    if isinstance(point_epilogue, _PlanPointEpi):
        if plan is not None and plan is not point_epilogue.plan:
            raise ValueError("point epilogue was built for another plan")
        return point_epilogue.plan
    if plan is not None:
        if point_epilogue is not None:
            raise ValueError("a plan call needs a point epilogue made by Plan.make_point_epilogue")
        return plan
    auto = _PLAN["plan"]
    if auto is None or point_epilogue is not None or pairs < _PLAN["min_pairs"]:
        return None
    if device_index is not None and (device_index not in auto.devices or not auto.peer_access):
        return None
    return auto


# ----------------------------------------------------------------------------------------
# flat (unstructured) entry points -- the reference signatures
# ----------------------------------------------------------------------------------------
def _flat(cov_samples, z_1, z_2, pos, vec, sf=None, epilogue=None, point_epilogue=None, plan=None):
    lib = _lib.load()
    epi = _epi_ref(epilogue)
    if sf is not None and epi is not None:
        raise ValueError("summate_fourier has no fused epilogue")
    if point_epilogue is not None and (vec or sf is not None):
        raise ValueError("the per-point epilogue applies to scalar fields only")
    if any(_is_cuda_tensor(x) for x in (cov_samples, z_1, z_2, pos)):
        return _flat_device(lib, cov_samples, z_1, z_2, pos, vec, sf, epi, point_epilogue, plan)
    cov = np.ascontiguousarray(_as_f64(cov_samples, "cov_samples"))
    z1 = np.ascontiguousarray(_as_f64(z_1, "z_1"))
    z2 = np.ascontiguousarray(_as_f64(z_2, "z_2"))
    p = _as_f64(pos, "pos")
    _check_modes(cov, z1, z2)
    if p.ndim != 2 or p.shape[0] != cov.shape[0]:
        raise ValueError("pos must have shape (dim, n) with the same dim as cov_samples")
    dim, n_modes = cov.shape
    n = p.shape[1]
    p, ld = _rows_contiguous(p)
    plan = None if sf is not None else _pick_plan(plan, float(n) * n_modes, point_epilogue)
    if plan is not None:
        out = _empty_host((dim, n) if vec else (n,))
        pe = point_epilogue.ref(n) if point_epilogue is not None else None
        rc = lib.gsb_plan_summate(plan.handle, _ptr(cov), _ptr(z1), _ptr(z2), _ptr(p), ld, dim, n_modes, n, _ptr(out),
                                  max(n, 1), int(vec), epi, pe, _lib.MEM_HOST, 0, None)
        _lib.check(rc, "plan_summate")
        return out
    if vec:
        out = _empty_host((dim, n))
        rc = lib.gsb_summate_incompr_ex(_ptr(cov), _ptr(z1), _ptr(z2), _ptr(p), ld, dim, n_modes, n,
                                        _ptr(out), max(n, 1), epi, _lib.MEM_HOST, get_device(), None)
        _lib.check(rc, "summate_incompr")
    elif sf is not None:
        f = np.ascontiguousarray(_as_f64(sf, "spectrum_factor"))
        if f.shape != z1.shape:
            raise ValueError("spectrum_factor must have shape (mode_no,)")
        out = _empty_host((n,))
        rc = lib.gsb_summate_fourier(_ptr(f), _ptr(cov), _ptr(z1), _ptr(z2), _ptr(p), ld, dim,
                                     n_modes, n, _ptr(out), _lib.MEM_HOST, get_device(), None)
        _lib.check(rc, "summate_fourier")
    elif point_epilogue is not None:
        device = point_epilogue.device.index if point_epilogue.device is not None else get_device()
        out = _empty_host((n,))
        rc = lib.gsb_summate_pp(_ptr(cov), _ptr(z1), _ptr(z2), _ptr(p), ld, dim, n_modes, n, _ptr(out), epi,
                                point_epilogue.ref(n, device), _lib.MEM_HOST, device, None)
        _lib.check(rc, "summate")
    else:
        out = _empty_host((n,))
        rc = lib.gsb_summate_ex(_ptr(cov), _ptr(z1), _ptr(z2), _ptr(p), ld, dim, n_modes, n,
                                _ptr(out), epi, _lib.MEM_HOST, get_device(), None)
        _lib.check(rc, "summate")
    return out


def _flat_device(lib, cov_samples, z_1, z_2, pos, vec, sf=None, epi=None, pepi=None, plan=None):
    torch = _torch()
    dev = next(x.device for x in (pos, cov_samples, z_1, z_2) if _is_cuda_tensor(x))

    def prep(x):
        return torch.as_tensor(x, dtype=torch.float64, device=dev).contiguous()

    cov, z1, z2 = prep(cov_samples), prep(z_1), prep(z_2)
    p = torch.as_tensor(pos, dtype=torch.float64, device=dev)
    ld = None
    if p.ndim == 2 and p.shape[1] > 0 and p.stride(1) == 1 and p.stride(0) >= p.shape[1]:
        ld = p.stride(0)  # row-strided view (a point range of a bigger array): no copy
    else:
        p = p.contiguous()
    if cov.ndim != 2 or z1.ndim != 1 or z2.ndim != 1 or z1.shape[0] != cov.shape[1] \
            or z2.shape[0] != cov.shape[1]:
        raise ValueError("cov_samples must be (dim, mode_no); z_1, z_2 must be (mode_no,)")
    if p.ndim != 2 or p.shape[0] != cov.shape[0]:
        raise ValueError("pos must have shape (dim, n) with the same dim as cov_samples")
    dim, n_modes = cov.shape
    n = p.shape[1]
    ld = ld or max(n, 1)
    stream = torch.cuda.current_stream(dev).cuda_stream
    plan = None if sf is not None else _pick_plan(plan, float(n) * n_modes, pepi, dev.index)
    if plan is not None:
        out = torch.empty((dim, n) if vec else (n,), dtype=torch.float64, device=dev)
        pe = pepi.ref(n) if pepi is not None else None
        rc = lib.gsb_plan_summate(plan.handle, cov.data_ptr(), z1.data_ptr(), z2.data_ptr(), p.data_ptr(), ld, dim,
                                  n_modes, n, out.data_ptr(), max(n, 1), int(vec), epi, pe, _lib.MEM_DEVICE, dev.index,
                                  stream)
        _lib.check(rc, "plan_summate")
        return out
    if vec:
        out = torch.empty((dim, n), dtype=torch.float64, device=dev)
        rc = lib.gsb_summate_incompr_ex(cov.data_ptr(), z1.data_ptr(), z2.data_ptr(), p.data_ptr(),
                                        ld, dim, n_modes, n, out.data_ptr(), max(n, 1), epi,
                                        _lib.MEM_DEVICE, dev.index, stream)
        _lib.check(rc, "summate_incompr")
    elif sf is not None:
        f = prep(sf)
        if tuple(f.shape) != tuple(z1.shape):
            raise ValueError("spectrum_factor must have shape (mode_no,)")
        out = torch.empty((n,), dtype=torch.float64, device=dev)
        rc = lib.gsb_summate_fourier(f.data_ptr(), cov.data_ptr(), z1.data_ptr(), z2.data_ptr(),
                                     p.data_ptr(), ld, dim, n_modes, n, out.data_ptr(),
                                     _lib.MEM_DEVICE, dev.index, stream)
        _lib.check(rc, "summate_fourier")
    elif pepi is not None:
        out = torch.empty((n,), dtype=torch.float64, device=dev)
        rc = lib.gsb_summate_pp(cov.data_ptr(), z1.data_ptr(), z2.data_ptr(), p.data_ptr(), ld, dim, n_modes, n,
                                out.data_ptr(), epi, pepi.ref(n, dev.index), _lib.MEM_DEVICE, dev.index, stream)
        _lib.check(rc, "summate")
    else:
        out = torch.empty((n,), dtype=torch.float64, device=dev)
        rc = lib.gsb_summate_ex(cov.data_ptr(), z1.data_ptr(), z2.data_ptr(), p.data_ptr(), ld,
                                dim, n_modes, n, out.data_ptr(), epi, _lib.MEM_DEVICE, dev.index,
                                stream)
        _lib.check(rc, "summate")
    return out


def summate(cov_samples, z_1, z_2, pos, num_threads=None, *, epilogue=None, point_epilogue=None):
    """B200 replacement of the native ``summate`` (generator.py:42-48, math :193-199).

    ``epilogue`` / ``point_epilogue`` (keyword only, not in the reference): see :func:`make_epilogue`
    and :func:`make_point_epilogue`.
    """
    return _flat(cov_samples, z_1, z_2, pos, vec=False, epilogue=epilogue, point_epilogue=point_epilogue)


def summate_incompr(cov_samples, z_1, z_2, pos, num_threads=None, *, epilogue=None):
    """B200 replacement of the native ``summate_incompr`` (generator.py:51-64, math :479-495)."""
    return _flat(cov_samples, z_1, z_2, pos, vec=True, epilogue=epilogue)


def summate_fourier(spectrum_factor, modes, z_1, z_2, pos, num_threads=None):
    """B200 replacement of the native ``summate_fourier`` (generator.py:67-75, 685-692)."""
    return _flat(modes, z_1, z_2, pos, vec=False, sf=spectrum_factor)


def summate_fourier_structured(spectrum_factor, modes, z_1, z_2, axes, matrix=None):
    """``summate_fourier`` on the mesh spanned by ``axes`` (host arrays), cf. :func:`summate_structured`."""
    lib = _lib.load()
    axes = [np.ascontiguousarray(_as_f64(a, "axes")).reshape(-1) for a in axes]
    dim = len(axes)
    cov = np.ascontiguousarray(_as_f64(modes, "modes"))
    z1 = np.ascontiguousarray(_as_f64(z_1, "z_1"))
    z2 = np.ascontiguousarray(_as_f64(z_2, "z_2"))
    f = np.ascontiguousarray(_as_f64(spectrum_factor, "spectrum_factor"))
    _check_modes(cov, z1, z2)
    if cov.shape[0] != dim or f.shape != z1.shape:
        raise ValueError("modes (dim, N), spectrum_factor (N,), len(axes) == dim")
    lens = np.array([a.shape[0] for a in axes], dtype=np.int64)
    cat = np.ascontiguousarray(np.concatenate(axes))
    mat_ptr = None
    if matrix is not None:
        mat = np.ascontiguousarray(_as_f64(matrix, "matrix"))
        if mat.shape != (dim, dim):
            raise ValueError("matrix must have shape (dim, dim)")
        mat_ptr = _ptr(mat)
    out = _empty_host(tuple(int(v) for v in lens))
    rc = lib.gsb_summate_fourier_structured(_ptr(f), _ptr(cov), _ptr(z1), _ptr(z2), _ptr(cat),
                                            lens.ctypes.data_as(_lib._c_int64_p), mat_ptr, dim,
                                            cov.shape[1], _ptr(out), _lib.MEM_HOST, get_device(), None)
    _lib.check(rc, "summate_fourier_structured")
    return out


# ----------------------------------------------------------------------------------------
# structured (rectilinear mesh) entry points -- the side channel for mesh_type="structured"
# ----------------------------------------------------------------------------------------
def _structured(cov_samples, z_1, z_2, axes, matrix, vec, epilogue=None, point_epilogue=None, plan=None):
    lib = _lib.load()
    epi = _epi_ref(epilogue)
    axes = list(axes)
    dim = len(axes)
    if point_epilogue is not None and vec:
        raise ValueError("the per-point epilogue applies to scalar fields only")
    if any(_is_cuda_tensor(x) for x in (cov_samples, z_1, z_2, *axes)):
        return _structured_device(lib, cov_samples, z_1, z_2, axes, matrix, vec, epi, point_epilogue, plan)
    cov = np.ascontiguousarray(_as_f64(cov_samples, "cov_samples"))
    z1 = np.ascontiguousarray(_as_f64(z_1, "z_1"))
    z2 = np.ascontiguousarray(_as_f64(z_2, "z_2"))
    batched = cov.ndim == 3
    if not batched:
        _check_modes(cov, z1, z2)
        cov3, z1b, z2b = cov[None], z1[None], z2[None]
    else:
        if z1.shape != (cov.shape[0], cov.shape[2]) or z2.shape != z1.shape:
            raise ValueError("batched call: cov_samples (B, dim, N), z_1 / z_2 (B, N)")
        cov3, z1b, z2b = cov, z1, z2
    if cov3.shape[1] != dim:
        raise ValueError("number of axes must equal the dim of cov_samples")
    ax = [np.ascontiguousarray(_as_f64(a, "axes")).reshape(-1) for a in axes]
    lens = (np.array([a.shape[0] for a in ax], dtype=np.int64))
    cat = np.ascontiguousarray(np.concatenate(ax)) if dim else np.empty(0)
    mat_ptr = None
    if matrix is not None:
        mat = np.ascontiguousarray(_as_f64(matrix, "matrix"))
        if mat.shape != (dim, dim):
            raise ValueError("matrix must have shape (dim, dim)")
        mat_ptr = _ptr(mat)
    shape = tuple(int(v) for v in lens)
    n_batch, _, n_modes = cov3.shape
    full = ((n_batch,) if batched else ()) + ((dim,) if vec else ()) + shape
    out = _empty_host(full)
    n_pts = int(np.prod(lens)) if dim else 0
    plan = _pick_plan(plan, float(n_pts) * n_modes * n_batch, point_epilogue)
    if plan is not None:
        pe = point_epilogue.ref(n_pts) if point_epilogue is not None else None
        rc = lib.gsb_plan_summate_structured(plan.handle, _ptr(cov3), _ptr(z1b), _ptr(z2b), _ptr(cat),
                                             lens.ctypes.data_as(_lib._c_int64_p), mat_ptr, dim, n_modes, n_batch,
                                             _ptr(out), int(vec), epi, pe, _lib.MEM_HOST, 0, None)
        _lib.check(rc, "plan_summate_structured")
        return out
    if point_epilogue is not None:
        device = point_epilogue.device.index if point_epilogue.device is not None else get_device()
        rc = lib.gsb_summate_structured_pp(_ptr(cov3), _ptr(z1b), _ptr(z2b), _ptr(cat),
                                           lens.ctypes.data_as(_lib._c_int64_p), mat_ptr, dim, n_modes, n_batch,
                                           _ptr(out), epi, point_epilogue.ref(int(np.prod(lens)), device),
                                           _lib.MEM_HOST, device, None)
        _lib.check(rc, "summate_structured")
        return out
    fn = lib.gsb_summate_incompr_structured_ex if vec else lib.gsb_summate_structured_ex
    rc = fn(_ptr(cov3), _ptr(z1b), _ptr(z2b), _ptr(cat),
            lens.ctypes.data_as(_lib._c_int64_p), mat_ptr, dim, n_modes, n_batch, _ptr(out), epi,
            _lib.MEM_HOST, get_device(), None)
    _lib.check(rc, "summate_incompr_structured" if vec else "summate_structured")
    return out


def _structured_device(lib, cov_samples, z_1, z_2, axes, matrix, vec, epi=None, pepi=None, plan=None):
    torch = _torch()
    dev = next(x.device for x in (cov_samples, z_1, z_2, *axes) if _is_cuda_tensor(x))

    def prep(x):
        return torch.as_tensor(x, dtype=torch.float64, device=dev).contiguous()

    cov, z1, z2 = prep(cov_samples), prep(z_1), prep(z_2)
    batched = cov.ndim == 3
    if not batched:
        cov, z1, z2 = cov[None], z1[None], z2[None]
    dim = len(axes)
    if cov.ndim != 3 or cov.shape[1] != dim or tuple(z1.shape) != (cov.shape[0], cov.shape[2]) \
            or z2.shape != z1.shape:
        raise ValueError("cov_samples (B, dim, N) / (dim, N); z_1, z_2 (B, N) / (N,); len(axes) == dim")
    ax = [prep(a).reshape(-1) for a in axes]
    lens = np.array([int(a.shape[0]) for a in ax], dtype=np.int64)
    cat = torch.cat(ax) if dim else torch.empty(0, dtype=torch.float64, device=dev)
    mat_ptr = None
    if matrix is not None:
        if _is_cuda_tensor(matrix):
            matrix = matrix.detach().cpu().numpy()
        mat = np.ascontiguousarray(_as_f64(matrix, "matrix"))
        if mat.shape != (dim, dim):
            raise ValueError("matrix must have shape (dim, dim)")
        mat_ptr = _ptr(mat)  # host pointer: the ABI reads the tiny matrix on the host
    shape = tuple(int(v) for v in lens)
    n_batch, _, n_modes = cov.shape
    full = ((n_batch,) if batched else ()) + ((dim,) if vec else ()) + shape
    out = torch.empty(full, dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    n_pts = int(np.prod(lens)) if dim else 0
    plan = _pick_plan(plan, float(n_pts) * n_modes * n_batch, pepi, dev.index)
    if plan is not None:
        pe = pepi.ref(n_pts) if pepi is not None else None
        rc = lib.gsb_plan_summate_structured(plan.handle, cov.data_ptr(), z1.data_ptr(), z2.data_ptr(), cat.data_ptr(),
                                             lens.ctypes.data_as(_lib._c_int64_p), mat_ptr, dim, n_modes, n_batch,
                                             out.data_ptr(), int(vec), epi, pe, _lib.MEM_DEVICE, dev.index, stream)
        _lib.check(rc, "plan_summate_structured")
        return out
    if pepi is not None:
        rc = lib.gsb_summate_structured_pp(cov.data_ptr(), z1.data_ptr(), z2.data_ptr(), cat.data_ptr(),
                                           lens.ctypes.data_as(_lib._c_int64_p), mat_ptr, dim, n_modes, n_batch,
                                           out.data_ptr(), epi, pepi.ref(int(np.prod(lens)), dev.index),
                                           _lib.MEM_DEVICE, dev.index, stream)
        _lib.check(rc, "summate_structured")
        return out
    fn = lib.gsb_summate_incompr_structured_ex if vec else lib.gsb_summate_structured_ex
    rc = fn(cov.data_ptr(), z1.data_ptr(), z2.data_ptr(), cat.data_ptr(),
            lens.ctypes.data_as(_lib._c_int64_p), mat_ptr, dim, n_modes, n_batch, out.data_ptr(),
            epi, _lib.MEM_DEVICE, dev.index, stream)
    _lib.check(rc, "summate_incompr_structured" if vec else "summate_structured")
    return out


def structured_slab_into(cov_samples, z_1, z_2, axes, matrix, lo, hi, out_ptr, *, incompr=False, epilogue=None):
    """Evaluate the entries ``[lo, hi)`` of axis 0 of the mesh ``axes`` and store them in place into the FULL
    field at the raw device address ``out_ptr`` (``gsb_summate_structured_slab``, device route).  Inputs are CUDA
    tensors on the calling rank's device; ``out_ptr`` may be memory of another GPU mapped into this process (peer
    access / :func:`gstools_b200.dist.open_peer_field`): the kernel then stores over NVLink and the sum needs no
    gather afterwards.  Work is enqueued on the current torch stream; nothing is returned."""
    lib = _lib.load()
    torch = _torch()
    dev = next(x.device for x in (cov_samples, z_1, z_2, *axes) if _is_cuda_tensor(x))

    def prep(x):
        return torch.as_tensor(x, dtype=torch.float64, device=dev).contiguous()

    cov, z1, z2 = prep(cov_samples), prep(z_1), prep(z_2)
    if cov.ndim == 2:
        cov, z1, z2 = cov[None], z1[None], z2[None]
    ax = [prep(a).reshape(-1) for a in axes]
    dim = len(ax)
    if cov.ndim != 3 or cov.shape[1] != dim or tuple(z1.shape) != (cov.shape[0], cov.shape[2]) or z2.shape != z1.shape:
        raise ValueError("cov_samples (B, dim, N) / (dim, N); z_1, z_2 (B, N) / (N,); len(axes) == dim")
    lens = np.array([int(a.shape[0]) for a in ax], dtype=np.int64)
    cat = torch.cat(ax)
    mat_ptr = None
    if matrix is not None:
        mat = np.ascontiguousarray(_as_f64(matrix, "matrix"))
        if mat.shape != (dim, dim):
            raise ValueError("matrix must have shape (dim, dim)")
        mat_ptr = _ptr(mat)
    stream = torch.cuda.current_stream(dev).cuda_stream
    rc = lib.gsb_summate_structured_slab(cov.data_ptr(), z1.data_ptr(), z2.data_ptr(), cat.data_ptr(),
                                         lens.ctypes.data_as(_lib._c_int64_p), mat_ptr, dim, cov.shape[2], cov.shape[0],
                                         int(lo), int(hi), int(out_ptr), int(bool(incompr)), _epi_ref(epilogue), None,
                                         _lib.MEM_DEVICE, dev.index, stream)
    _lib.check(rc, "summate_structured_slab")


def summate_structured(cov_samples, z_1, z_2, axes, matrix=None, *, epilogue=None, point_epilogue=None):
    """``summate`` on the mesh spanned by ``axes`` without the flat position array.

    Equals ``summate(cov_samples, z_1, z_2, matrix @ generate_grid(axes)).reshape(shape)``
    (reference: field/base.py:289-297, tools/geometric.py:340-356, covmodel/base.py:572-582).
    ``cov_samples`` may carry a leading batch axis (ensembles of mode sets on one mesh).
    """
    return _structured(cov_samples, z_1, z_2, axes, matrix, vec=False, epilogue=epilogue,
                       point_epilogue=point_epilogue)


def summate_incompr_structured(cov_samples, z_1, z_2, axes, matrix=None, *, epilogue=None):
    """Vector-field variant of :func:`summate_structured`; returns ``(dim,) + shape``."""
    return _structured(cov_samples, z_1, z_2, axes, matrix, vec=True, epilogue=epilogue)


# ----------------------------------------------------------------------------------------
# kriging evaluation -- the reference signatures (krige/base.py:42-61)
# ----------------------------------------------------------------------------------------
def _krige(krig_mat, krig_vecs, cond, want_var):
    lib = _lib.load()
    name = "calc_field_krige_and_variance" if want_var else "calc_field_krige"
    if any(_is_cuda_tensor(x) for x in (krig_mat, krig_vecs, cond)):
        torch = _torch()
        dev = next(x.device for x in (krig_vecs, krig_mat, cond) if _is_cuda_tensor(x))
        mat = torch.as_tensor(krig_mat, dtype=torch.float64, device=dev).contiguous()
        c = torch.as_tensor(cond, dtype=torch.float64, device=dev).contiguous()
        kv = torch.as_tensor(krig_vecs, dtype=torch.float64, device=dev)
        if kv.ndim == 2 and kv.shape[1] > 0 and kv.stride(1) == 1 and kv.stride(0) >= kv.shape[1]:
            ld = kv.stride(0)
        else:
            kv = kv.contiguous()
            ld = max(kv.shape[1], 1) if kv.ndim == 2 else 1
        if mat.ndim != 2 or mat.shape[0] != mat.shape[1] or kv.ndim != 2 or kv.shape[0] != mat.shape[0] \
                or tuple(c.shape) != (mat.shape[0],):
            raise ValueError("krig_mat (K, K), krig_vecs (K, n), cond (K,)")
        size, n = mat.shape[0], kv.shape[1]
        field = torch.empty(n, dtype=torch.float64, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        if want_var:
            error = torch.empty(n, dtype=torch.float64, device=dev)
            rc = lib.gsb_calc_field_krige_and_variance(mat.data_ptr(), kv.data_ptr(), ld, c.data_ptr(),
                                                       size, n, field.data_ptr(), error.data_ptr(),
                                                       _lib.MEM_DEVICE, dev.index, stream)
            _lib.check(rc, name)
            return field, error
        rc = lib.gsb_calc_field_krige(mat.data_ptr(), kv.data_ptr(), ld, c.data_ptr(), size, n,
                                      field.data_ptr(), _lib.MEM_DEVICE, dev.index, stream)
        _lib.check(rc, name)
        return field
    mat = np.ascontiguousarray(_as_f64(krig_mat, "krig_mat"))
    c = np.ascontiguousarray(_as_f64(cond, "cond"))
    kv = _as_f64(krig_vecs, "krig_vecs")
    if mat.ndim != 2 or mat.shape[0] != mat.shape[1] or kv.ndim != 2 or kv.shape[0] != mat.shape[0] \
            or c.shape != (mat.shape[0],):
        raise ValueError("krig_mat (K, K), krig_vecs (K, n), cond (K,)")
    size, n = mat.shape[0], kv.shape[1]
    kv, ld = _rows_contiguous(kv)
    field = _empty_host((n,))
    if want_var:
        error = _empty_host((n,))
        rc = lib.gsb_calc_field_krige_and_variance(_ptr(mat), _ptr(kv), ld, _ptr(c), size, n,
                                                   _ptr(field), _ptr(error), _lib.MEM_HOST,
                                                   get_device(), None)
        _lib.check(rc, name)
        return field, error
    rc = lib.gsb_calc_field_krige(_ptr(mat), _ptr(kv), ld, _ptr(c), size, n, _ptr(field),
                                  _lib.MEM_HOST, get_device(), None)
    _lib.check(rc, name)
    return field


def calc_field_krige_and_variance(krig_mat, krig_vecs, cond, num_threads=None):
    """B200 replacement of the native ``calc_field_krige_and_variance`` (krige/base.py:51-61):
    ``(field, error)`` with ``field = cond @ (krig_mat @ krig_vecs)`` and
    ``error[k] = krig_vecs[:, k] @ krig_mat @ krig_vecs[:, k]``."""
    return _krige(krig_mat, krig_vecs, cond, True)


def calc_field_krige(krig_mat, krig_vecs, cond, num_threads=None):
    """B200 replacement of the native ``calc_field_krige`` (krige/base.py:42-49)."""
    return _krige(krig_mat, krig_vecs, cond, False)


def cov_model_spec(kind, var, len_rescaled, sill=None, param=0.0, exact=False):
    """``gsb_cov_model``: ``var * cor(r / len_rescaled)`` with ``cor`` the closed form of model ``kind``
    (one of ``_lib.COV_TYPES``; src/gstools/covmodel/models.py), ``sill`` at r ~ 0 when ``exact``."""
    if kind not in _lib.COV_TYPES:
        raise ValueError(f"covariance model '{kind}' has no device implementation: {sorted(_lib.COV_TYPES)}")
    spec = _lib.CovModelSpec()
    spec.type = _lib.COV_TYPES[kind]
    spec.exact = int(bool(exact))
    spec.var = float(var)
    spec.len_rescaled = float(len_rescaled)
    spec.sill = float(var if sill is None else sill)
    spec.param = float(param)
    return spec


def _krige_evaluate_device(lib, model, krig_mat, cond, cond_pos, pos, axes, matrix, unbiased, tail_rows,
                           return_var, plan=None):
    """Device-resident variant: CUDA tensors in, CUDA tensors out, work enqueued on the current stream."""
    torch = _torch()
    cands = [krig_mat, cond, cond_pos, pos, tail_rows] + (list(axes) if axes is not None else [])
    dev = next(x.device for x in cands if _is_cuda_tensor(x))

    def prep(x):
        return torch.as_tensor(x, dtype=torch.float64, device=dev).contiguous()

    mat, c, cp = prep(krig_mat), prep(cond), prep(cond_pos)
    if mat.ndim != 2 or mat.shape[0] != mat.shape[1] or tuple(c.shape) != (mat.shape[0],) or cp.ndim != 2:
        raise ValueError("krig_mat (K, K), cond (K,), cond_pos (dim, cond_no)")
    size, (dim, cond_no) = mat.shape[0], cp.shape
    n_tail = size - cond_no - int(bool(unbiased))
    if n_tail < 0:
        raise ValueError("cond_no + unbiased exceeds the kriging system size")
    if (pos is None) == (axes is None):
        raise ValueError("give either pos (dim, n) or axes")
    if axes is not None:
        ax = [prep(a).reshape(-1) for a in axes]
        if len(ax) != dim:
            raise ValueError("number of axes must equal the dim of cond_pos")
        lens = np.array([int(a.shape[0]) for a in ax], dtype=np.int64)
        shape, n = tuple(int(v) for v in lens), int(np.prod(lens))
        cat = torch.cat(ax)
        mat_ptr = None
        if matrix is not None:
            if _is_cuda_tensor(matrix):
                matrix = matrix.detach().cpu().numpy()
            m = np.ascontiguousarray(_as_f64(matrix, "matrix"))
            if m.shape != (dim, dim):
                raise ValueError("matrix must have shape (dim, dim)")
            mat_ptr = _ptr(m)            # host pointer: the ABI reads the tiny matrix on the host
    else:
        p = prep(pos)
        if p.ndim != 2 or p.shape[0] != dim:
            raise ValueError("pos must have shape (dim, n) with the dim of cond_pos")
        n, shape = p.shape[1], (p.shape[1],)
    tail_ptr, tail_ld = None, max(n, 1)
    if n_tail > 0:
        if tail_rows is None:
            raise ValueError(f"{n_tail} drift rows expected in tail_rows")
        tail = prep(tail_rows).reshape(n_tail, -1)
        if tail.shape[1] != n:
            raise ValueError("tail_rows must have shape (krige_size - cond_no - unbiased, n)")
        tail_ptr = tail.data_ptr()
    field = torch.empty(n, dtype=torch.float64, device=dev)
    error = torch.empty(n, dtype=torch.float64, device=dev) if return_var else None
    err_ptr = error.data_ptr() if return_var else None
    stream = torch.cuda.current_stream(dev).cuda_stream
    plan = _pick_plan(plan, float(n) * size * size / 4.0, None, dev.index)
    if plan is not None:
        rc = lib.gsb_plan_krige_evaluate(plan.handle, ctypes.byref(model), mat.data_ptr(), c.data_ptr(), size,
                                         cp.data_ptr(), cond_no, dim, None if axes is not None else p.data_ptr(),
                                         max(n, 1), n, cat.data_ptr() if axes is not None else None,
                                         lens.ctypes.data_as(_lib._c_int64_p) if axes is not None else None,
                                         mat_ptr if axes is not None else None, int(bool(unbiased)), tail_ptr, tail_ld,
                                         field.data_ptr(), err_ptr, _lib.MEM_DEVICE, dev.index, stream)
    elif axes is not None:
        rc = lib.gsb_krige_evaluate_structured(ctypes.byref(model), mat.data_ptr(), c.data_ptr(), size,
                                               cp.data_ptr(), cond_no, dim, cat.data_ptr(),
                                               lens.ctypes.data_as(_lib._c_int64_p), mat_ptr,
                                               int(bool(unbiased)), tail_ptr, tail_ld, field.data_ptr(), err_ptr,
                                               _lib.MEM_DEVICE, dev.index, stream)
    else:
        rc = lib.gsb_krige_evaluate(ctypes.byref(model), mat.data_ptr(), c.data_ptr(), size, cp.data_ptr(),
                                    cond_no, dim, p.data_ptr(), max(n, 1), n, int(bool(unbiased)), tail_ptr,
                                    tail_ld, field.data_ptr(), err_ptr, _lib.MEM_DEVICE, dev.index, stream)
    _lib.check(rc, "krige_evaluate")
    field = field.reshape(shape)
    return (field, error.reshape(shape)) if return_var else field


def krige_evaluate(model, krig_mat, cond, cond_pos, pos=None, axes=None, matrix=None, unbiased=True,
                   tail_rows=None, return_var=True, plan=None):
    """The evaluation loop of ``Krige.__call__`` (krige/base.py:278-294) on the device.

    The right-hand sides of ``Krige._get_krige_vecs`` (base.py:359-388) -- ``model`` evaluated at
    the distances between ``cond_pos`` (dim, cond_no) and the evaluation points, the row of ones of
    an unbiased system, then ``tail_rows`` (drift rows, evaluated by the caller) -- are generated
    on the GPU and contracted with ``krig_mat`` / ``cond`` there.  Evaluation points: ``pos``
    (dim, n), isometrised, or the mesh ``axes`` (+ isometrisation ``matrix``).  Host arrays in,
    host arrays out -- or CUDA tensors in, CUDA tensors out (nothing is copied, work is enqueued on the
    current torch stream).  Returns ``(field, error)`` or ``field`` (``return_var=False``); for a mesh
    the results have the mesh shape.  ``plan`` (or the process-wide plan of :func:`use_devices` for big systems):
    the points are dealt out to the GPUs of the plan -- slabs along axis 0 of a mesh, ranges of a flat point set.
    """
    lib = _lib.load()
    if not isinstance(model, _lib.CovModelSpec):
        model = cov_model_spec(**model)
    tensors = [krig_mat, cond, cond_pos, pos, tail_rows] + (list(axes) if axes is not None else [])
    if any(_is_cuda_tensor(x) for x in tensors):
        return _krige_evaluate_device(lib, model, krig_mat, cond, cond_pos, pos, axes, matrix, unbiased,
                                      tail_rows, return_var, plan)
    mat = np.ascontiguousarray(_as_f64(krig_mat, "krig_mat"))
    c = np.ascontiguousarray(_as_f64(cond, "cond"))
    cp = np.ascontiguousarray(_as_f64(cond_pos, "cond_pos"))
    if mat.ndim != 2 or mat.shape[0] != mat.shape[1] or c.shape != (mat.shape[0],) or cp.ndim != 2:
        raise ValueError("krig_mat (K, K), cond (K,), cond_pos (dim, cond_no)")
    size, (dim, cond_no) = mat.shape[0], cp.shape
    n_tail = size - cond_no - int(bool(unbiased))
    if n_tail < 0:
        raise ValueError("cond_no + unbiased exceeds the kriging system size")
    if (pos is None) == (axes is None):
        raise ValueError("give either pos (dim, n) or axes")
    if axes is not None:
        ax = [np.ascontiguousarray(_as_f64(a, "axes")).reshape(-1) for a in axes]
        if len(ax) != dim:
            raise ValueError("number of axes must equal the dim of cond_pos")
        lens = np.array([a.shape[0] for a in ax], dtype=np.int64)
        shape = tuple(int(v) for v in lens)
        n = int(np.prod(lens))
        cat = np.ascontiguousarray(np.concatenate(ax))
        mat_ptr = None
        if matrix is not None:
            m = np.ascontiguousarray(_as_f64(matrix, "matrix"))
            if m.shape != (dim, dim):
                raise ValueError("matrix must have shape (dim, dim)")
            mat_ptr = _ptr(m)
    else:
        p = _as_f64(pos, "pos")
        if p.ndim != 2 or p.shape[0] != dim:
            raise ValueError("pos must have shape (dim, n) with the dim of cond_pos")
        n = p.shape[1]
        shape = (n,)
        p, ld = _rows_contiguous(p)
    tail_ptr, tail_ld = None, max(n, 1)
    if n_tail > 0:
        if tail_rows is None:
            raise ValueError(f"{n_tail} drift rows expected in tail_rows")
        tail = _as_f64(tail_rows, "tail_rows").reshape(n_tail, -1)
        if tail.shape[1] != n:
            raise ValueError("tail_rows must have shape (krige_size - cond_no - unbiased, n)")
        tail, tail_ld = _rows_contiguous(tail)
        tail_ptr = _ptr(tail)
    field = _empty_host((n,))
    error = _empty_host((n,)) if return_var else None
    err_ptr = _ptr(error) if return_var else None
    plan = _pick_plan(plan, float(n) * size * size / 4.0, None)
    if plan is not None:
        rc = lib.gsb_plan_krige_evaluate(plan.handle, ctypes.byref(model), _ptr(mat), _ptr(c), size, _ptr(cp), cond_no,
                                         dim, None if axes is not None else _ptr(p), ld if axes is None else 1, n,
                                         _ptr(cat) if axes is not None else None,
                                         lens.ctypes.data_as(_lib._c_int64_p) if axes is not None else None,
                                         mat_ptr if axes is not None else None, int(bool(unbiased)), tail_ptr, tail_ld,
                                         _ptr(field), err_ptr, _lib.MEM_HOST, 0, None)
    elif axes is not None:
        rc = lib.gsb_krige_evaluate_structured(ctypes.byref(model), _ptr(mat), _ptr(c), size, _ptr(cp),
                                               cond_no, dim, _ptr(cat), lens.ctypes.data_as(_lib._c_int64_p),
                                               mat_ptr, int(bool(unbiased)), tail_ptr, tail_ld, _ptr(field),
                                               err_ptr, _lib.MEM_HOST, get_device(), None)
    else:
        rc = lib.gsb_krige_evaluate(ctypes.byref(model), _ptr(mat), _ptr(c), size, _ptr(cp), cond_no, dim,
                                    _ptr(p), ld, n, int(bool(unbiased)), tail_ptr, tail_ld, _ptr(field),
                                    err_ptr, _lib.MEM_HOST, get_device(), None)
    _lib.check(rc, "krige_evaluate")
    field = field.reshape(shape)
    return (field, error.reshape(shape)) if return_var else field


def _mt_keys(burn_state, main_state):
    keys = []
    for st in (burn_state, main_state):
        if st[0] != "MT19937":
            raise ValueError("the legacy MT19937 state of numpy.random.RandomState is required")
        key = np.ascontiguousarray(st[1], dtype=np.uint32)
        if key.shape != (624,):
            raise ValueError("MT19937 key must have 624 words")
        keys.append(key)
    return keys


def sample_radii_mcmc(kind, dim, len_rescaled, nu, burn_state, main_state, init, burn_in, n_steps):
    """Native, stream-compatible ``emcee`` run of ``RNG.sample_ln_pdf`` (random/rng.py:77-101).

    ``burn_state`` / ``main_state`` are the ``numpy.random.RandomState.get_state()`` tuples (legacy
    MT19937) the reference hands to the burn-in and the production ``run_mcmc`` call, ``init`` the
    ``nwalkers`` initial positions.  Returns the production chain ``(n_steps, nwalkers)``.  Host only.
    ``kind``: a model name with a native log-pdf (``_lib.PDF_KINDS``), or a callable ``ln_pdf(r)`` -- the
    model's own vectorised ``ln_spectral_rad_pdf``, called with an ``(n, 1)`` array like emcee does -- for
    every other model (``gsb_sample_radii_mcmc_cb``).
    """
    keys = _mt_keys(burn_state, main_state)
    x0 = np.ascontiguousarray(_as_f64(init, "init")).reshape(-1)
    chain = np.empty((int(n_steps), x0.shape[0]), dtype=np.float64)
    lib = _lib.load()
    if callable(kind):
        ln_pdf, failure = kind, []

        def trampoline(r_ptr, n, out_ptr, _user):
            try:      # no exception may cross the C frames: remember it, abort the chain, re-raise below
                r = np.ctypeslib.as_array(r_ptr, shape=(n,))
                vals = np.asarray(ln_pdf(np.array(r).reshape(n, 1)), dtype=np.float64).reshape(-1)
                if vals.shape[0] != n:
                    raise ValueError("ln_pdf must return one value per radius")
                np.ctypeslib.as_array(out_ptr, shape=(n,))[:] = vals
                return 0
            except BaseException as exc:  # noqa: BLE001
                failure.append(exc)
                return 1

        cb = _lib.LN_PDF_FN(trampoline)
        rc = lib.gsb_sample_radii_mcmc_cb(cb, None, _ptr(keys[0]), int(burn_state[2]), _ptr(keys[1]),
                                          int(main_state[2]), _ptr(x0), x0.shape[0], int(burn_in), int(n_steps),
                                          _ptr(chain))
        if failure:
            raise failure[0]
        _lib.check(rc, "sample_radii_mcmc")
        return chain
    if kind not in _lib.PDF_KINDS:
        raise ValueError(f"no native log-pdf for model '{kind}': {sorted(_lib.PDF_KINDS)}")
    rc = lib.gsb_sample_radii_mcmc(_lib.PDF_KINDS[kind], int(dim), float(len_rescaled), float(nu),
                                   _ptr(keys[0]), int(burn_state[2]), _ptr(keys[1]), int(main_state[2]),
                                   _ptr(x0), x0.shape[0], int(burn_in), int(n_steps), _ptr(chain))
    _lib.check(rc, "sample_radii_mcmc")
    return chain


def sample_modes_batch(kind, dim, len_rescaled, nu, seeds, mode_no, nwalkers=50, burn_in=20, oversampling_factor=10,
                       num_threads=None):
    """Mode sets ``(cov_samples (S, dim, N), z_1 (S, N), z_2 (S, N))`` of ``RandMeth(model, mode_no=N, seed=s)`` for every
    seed ``s`` of ``seeds``, bit for bit what ``RandMeth.reset_seed`` (generator.py:346-387) draws one seed at a time --
    for models whose radii come from ``RNG.sample_ln_pdf`` and have a native log-pdf (``kind`` in ``_lib.PDF_KINDS``).
    The random streams (normal, uniform, the emcee chain, choice) run natively, one seed per task on ``num_threads``
    host threads (default: all cores); the sphere coordinates are finished with numpy on the whole batch."""
    if kind not in _lib.PDF_KINDS:
        raise ValueError(f"no native log-pdf for model '{kind}': {sorted(_lib.PDF_KINDS)}")
    if dim not in (1, 2, 3):
        raise ValueError("sample_sphere supports dim 1, 2 and 3 natively")
    seeds = np.ascontiguousarray(seeds, dtype=np.int64).reshape(-1)
    n_seeds, n = seeds.shape[0], int(mode_no)
    sample_size = int(max(burn_in, (n / nwalkers) * oversampling_factor))          # rng.py:78-83
    if num_threads is None:
        try:
            num_threads = len(os.sched_getaffinity(0))
        except AttributeError:
            num_threads = os.cpu_count() or 1
    z1, z2, ang1, rad = (np.empty((n_seeds, n)) for _ in range(4))
    ang2 = np.empty((n_seeds, n)) if dim == 3 else None
    rc = _lib.load().gsb_sample_modes_batch(_lib.PDF_KINDS[kind], int(dim), float(len_rescaled), float(nu), _ptr(seeds),
                                            n_seeds, n, int(nwalkers), int(burn_in), sample_size,
                                            1.0 / float(len_rescaled), 2 * np.pi, int(num_threads), _ptr(z1), _ptr(z2),
                                            _ptr(ang1), _ptr(ang2) if ang2 is not None else None, _ptr(rad))
    _lib.check(rc, "sample_modes_batch")
    coord = np.empty((n_seeds, dim, n))
    if dim == 1:                                                                     # rng.py:163-174
        coord[:, 0] = ang1
    elif dim == 2:
        coord[:, 0] = np.cos(ang1)
        coord[:, 1] = np.sin(ang1)
    else:
        coord[:, 0] = np.sqrt(1.0 - ang2**2) * np.cos(ang1)
        coord[:, 1] = np.sqrt(1.0 - ang2**2) * np.sin(ang1)
        coord[:, 2] = ang2
    return rad[:, None, :] * coord, z1, z2                                           # generator.py:387


def scale_shift_(field, scale, shift=0.0):
    """In-place ``field = scale*field + shift`` on a CUDA tensor (generator.py:269-270)."""
    if not _is_cuda_tensor(field):
        raise TypeError("scale_shift_ works on CUDA tensors; use numpy for host arrays")
    torch = _torch()
    if field.dtype != torch.float64 or not field.is_contiguous():
        raise ValueError("field must be a contiguous float64 CUDA tensor")
    stream = torch.cuda.current_stream(field.device).cuda_stream
    rc = _lib.load().gsb_scale_shift(field.data_ptr(), field.numel(), float(scale), float(shift),
                                     field.device.index, stream)
    _lib.check(rc, "scale_shift")
    return field
