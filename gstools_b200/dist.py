"""Multi-GPU sharding of the summation: one process per GPU, no traffic during the sum.

Every output element depends only on its own position and on the (tiny, replicated) mode set,
so the path shards into independent units:

* flat point sets       -> contiguous point ranges (:func:`shard_range`),
* structured meshes     -> slabs along axis 0, so each rank's output is one contiguous block of
                           the C-ordered field (reference layout: tools/geometric.py:340-356),
* ensembles of seeds    -> contiguous ranges of realisations (README.md:255-257 idiom).

NCCL (``torch.distributed``) is used only when the caller asks for the field as ONE array:
:func:`gather_field` all-gathers or gathers finished per-rank slabs straight into the destination
array; :func:`summate_structured_gathered` overlaps that gather with the sum -- either NCCL
send/recv of row pieces on a side stream while the next piece contracts (``mode="nccl"``), or no
collective at all: the destination field of rank ``dst`` is mapped into every rank (CUDA IPC) and
each rank's contraction kernel stores its slab there over NVLink (``mode="p2p"``).  The compute
functions are injectable so the partition + gather logic is testable on CPU with the gloo backend.
"""

from __future__ import annotations

import numpy as np

from . import backend

__all__ = [
    "shard_range",
    "summate_sharded",
    "summate_structured_sharded",
    "ensemble_sharded",
    "krige_evaluate_sharded",
    "gather_field",
    "piece_bounds",
    "summate_structured_gathered",
    "open_peer_field",
]


def _dist():
    import torch.distributed as dist

    return dist


def _rank_world(group=None):
    dist = _dist()
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_range(n: int, rank: int, world: int):
    """Contiguous, balanced ``[lo, hi)`` of ``n`` units for ``rank`` (sizes differ by <= 1)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("invalid rank/world")
    base, rem = divmod(int(n), world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def summate_sharded(cov_samples, z_1, z_2, pos, group=None, incompr=False, compute=None):
    """This rank's slice of ``summate(...)``: returns ``(local_out, (lo, hi))``.

    ``pos`` is the FULL ``(dim, n)`` array (host or device); only columns ``lo:hi`` are read.
    """
    rank, world = _rank_world(group)
    n = pos.shape[1]
    lo, hi = shard_range(n, rank, world)
    fn = compute or (backend.summate_incompr if incompr else backend.summate)
    return fn(cov_samples, z_1, z_2, pos[:, lo:hi]), (lo, hi)


def summate_structured_sharded(cov_samples, z_1, z_2, axes, matrix=None, group=None,
                               incompr=False, compute=None):
    """This rank's slab (range of axis 0) of the structured sum: ``(local_out, (lo, hi))``."""
    rank, world = _rank_world(group)
    axes = list(axes)
    lo, hi = shard_range(len(axes[0]), rank, world)
    fn = compute or (backend.summate_incompr_structured if incompr
                     else backend.summate_structured)
    local_axes = [axes[0][lo:hi]] + axes[1:]
    return fn(cov_samples, z_1, z_2, local_axes, matrix), (lo, hi)


def ensemble_sharded(mode_sets, evaluate, group=None):
    """Evaluate this rank's share of an ensemble.

    ``mode_sets`` is a sequence (one entry per realisation / seed); ``evaluate(entry)`` returns
    that realisation's field.  Returns ``(list_of_local_fields, (lo, hi))``.
    """
    rank, world = _rank_world(group)
    lo, hi = shard_range(len(mode_sets), rank, world)
    return [evaluate(mode_sets[i]) for i in range(lo, hi)], (lo, hi)


def krige_evaluate_sharded(model, krig_mat, cond, cond_pos, pos=None, axes=None, matrix=None,
                           unbiased=True, tail_rows=None, return_var=True, group=None, compute=None):
    """This rank's share of :func:`gstools_b200.krige_evaluate` (row f1): points are independent, the
    kriging system (``krig_mat``, ``cond``, ``cond_pos``) is replicated.  Flat points are split into
    contiguous ranges, meshes into slabs along axis 0; drift rows follow their points.
    Returns ``(local_result, (lo, hi))`` with ``local_result`` as ``krige_evaluate`` returns it."""
    rank, world = _rank_world(group)
    fn = compute or backend.krige_evaluate
    if (pos is None) == (axes is None):
        raise ValueError("give either pos (dim, n) or axes")
    if axes is not None:
        axes = [np.asarray(a, dtype=np.float64).reshape(-1) for a in axes]
        lo, hi = shard_range(len(axes[0]), rank, world)
        inner = int(np.prod([len(a) for a in axes[1:]])) if len(axes) > 1 else 1
        tail = None if tail_rows is None else np.asarray(tail_rows)[:, lo * inner:hi * inner]
        out = fn(model, krig_mat, cond, cond_pos, axes=[axes[0][lo:hi]] + axes[1:], matrix=matrix,
                 unbiased=unbiased, tail_rows=tail, return_var=return_var)
    else:
        lo, hi = shard_range(pos.shape[1], rank, world)
        tail = None if tail_rows is None else np.asarray(tail_rows)[:, lo:hi]
        out = fn(model, krig_mat, cond, cond_pos, pos=pos[:, lo:hi], unbiased=unbiased, tail_rows=tail,
                 return_var=return_var)
    return out, (lo, hi)


def gather_field(local, n_total: int, axis: int = 0, group=None, dst=None):
    """Assemble per-rank slabs into one array (the ONLY collective on this path).

    ``local`` is this rank's block (torch tensor on the rank's device, or numpy for gloo/CPU),
    split along ``axis`` according to :func:`shard_range`.  ``dst=None`` -> every rank gets the
    field (all-gather); ``dst=r`` -> only rank r does, others get ``None``.  Every slab travels
    straight into its place of the destination array (grouped send/recv, no padding, no concatenation).
    """
    import torch

    dist = _dist()
    rank, world = _rank_world(group)
    was_numpy = isinstance(local, np.ndarray)
    t = torch.from_numpy(np.ascontiguousarray(local)) if was_numpy else local
    if world == 1:
        return local
    t = t.movedim(axis, 0).contiguous()
    ranges = [shard_range(n_total, r, world) for r in range(world)]
    if t.shape[0] != ranges[rank][1] - ranges[rank][0]:
        raise ValueError("local block does not match shard_range(n_total, rank, world)")
    receiver = dst is None or rank == dst
    full = None
    ops = []
    if receiver:
        full = torch.empty((n_total,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        full[ranges[rank][0]:ranges[rank][1]].copy_(t)
        for r, (lo, hi) in enumerate(ranges):
            if r != rank and hi > lo:
                ops.append(dist.P2POp(dist.irecv, full[lo:hi], _global_rank(r, group), group))
    if t.shape[0] > 0:
        for q in (range(world) if dst is None else [dst]):
            if q != rank:
                ops.append(dist.P2POp(dist.isend, t, _global_rank(q, group), group))
    for req in (dist.batch_isend_irecv(ops) if ops else []):
        req.wait()
    if full is None:
        return None
    full = full.movedim(0, axis)
    return full.numpy() if was_numpy else full


def _global_rank(r, group):
    dist = _dist()
    return r if group is None else dist.get_global_rank(group, r)


def piece_bounds(n: int, pieces: int):
    """Cut ``n`` rows into at most ``pieces`` pieces of DECREASING size (1/2, 1/4, ..., the last two equal): the
    transfer of a piece hides behind the contraction of the next, and what stays exposed after the last
    contraction is only the smallest piece.  Returns the cut positions ``[0, ..., n]``."""
    n, pieces = int(n), max(1, min(int(pieces), int(n)))
    cuts, rem = [0], n
    for k in range(pieces - 1):
        take = min(max(1, rem // 2), rem - (pieces - 1 - k))
        cuts.append(cuts[-1] + take)
        rem -= take
    if n > 0 or not cuts[1:]:
        cuts.append(n)
    return cuts


_PEER = {}      # (shape, dst, group id) -> dict(full=tensor on dst | None, ptr=mapped address, base=..., device=...)


def open_peer_field(shape, dst=0, group=None, device=None):
    """The gathered field of rank ``dst`` -- a float64 CUDA tensor of ``shape`` -- mapped into EVERY rank of the
    group (``gsb_ipc_export`` / ``gsb_ipc_open``: CUDA IPC + peer access over NVLink; one node).  Returns
    ``(tensor or None, address)``: the tensor on ``dst``, and on every rank the address under which this rank's
    kernels can store into it.  Collective (handles are exchanged once per shape and cached; the buffer is REUSED by
    later calls with the same shape)."""
    import ctypes

    import torch

    from . import _lib

    dist = _dist()
    rank, world = _rank_world(group)
    key = (tuple(int(v) for v in shape), int(dst), id(group))
    hit = _PEER.get(key)
    if hit is not None:
        return hit["full"], hit["ptr"]
    lib = _lib.load()
    dev_index = torch.cuda.current_device() if device is None else int(device)
    payload = [None]
    full = None
    if rank == dst:
        full = torch.empty(key[0], dtype=torch.float64, device=torch.device("cuda", dev_index))
        handle = (ctypes.c_ubyte * 64)()
        off = ctypes.c_int64(0)
        _lib.check(lib.gsb_ipc_export(full.data_ptr(), dev_index, handle, ctypes.byref(off)), "ipc_export")
        payload = [(bytes(handle), int(off.value))]
    if world > 1:
        dist.broadcast_object_list(payload, src=_global_rank(dst, group), group=group)
    entry, problem = None, None
    if rank == dst:
        entry = dict(full=full, ptr=full.data_ptr(), base=None, device=dev_index)
    else:
        try:
            raw, off = payload[0]
            base = ctypes.c_void_p()
            buf = (ctypes.c_ubyte * 64).from_buffer_copy(raw)
            _lib.check(lib.gsb_ipc_open(buf, dev_index, ctypes.byref(base)), "ipc_open")
            entry = dict(full=None, ptr=int(base.value) + off, base=int(base.value), device=dev_index)
        except Exception as exc:  # noqa: BLE001  (reported on EVERY rank below: nobody may run ahead into a collective)
            problem = exc
    if world > 1:
        ok = torch.tensor([0.0 if problem is not None else 1.0], device=torch.device("cuda", dev_index))
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if float(ok.item()) < 1.0:
            if entry is not None and entry["base"] is not None:
                lib.gsb_ipc_close(entry["base"], dev_index)
            raise RuntimeError("open_peer_field: mapping the destination into every rank failed"
                               + (f" on this rank: {problem}" if problem is not None else " on another rank"))
    _PEER[key] = entry
    return entry["full"], entry["ptr"]


def close_peer_fields():
    """Unmap every field opened by :func:`open_peer_field` (call before the owner frees its tensors)."""
    from . import _lib

    lib = _lib.load()
    for entry in _PEER.values():
        if entry["base"] is not None:
            lib.gsb_ipc_close(entry["base"], entry["device"])
    _PEER.clear()


def summate_structured_gathered(cov_samples, z_1, z_2, axes, matrix=None, group=None, dst=0, incompr=False,
                                mode="nccl", pieces=1, compute=None, reserve_sms=0):
    """The structured sum, sharded into slabs along axis 0 over the ranks of ``group``, delivered as ONE array on
    rank ``dst`` (``dst=None``: on every rank; ``mode="nccl"`` only).  Other ranks get ``None``.

    ``mode="nccl"``: the finished slab travels as one grouped NCCL send/recv straight into its rows of the destination
    (``pieces=1``, the default).  With ``pieces > 1`` every rank cuts its slab into row pieces of decreasing size
    (:func:`piece_bounds`) and piece k travels on a side stream while piece k+1 contracts -- implemented and tested,
    but measured NOT to pay on B200: the contraction is a persistent kernel that owns every SM, so NCCL's kernels get
    no SM until it ends (profiles/r02_gather_variants_2gpu.log).
    ``mode="p2p"``: no collective on the data path at all -- :func:`open_peer_field` maps the destination into every
    rank and each rank's ONE contraction launch stores its slab there from the kernel's epilogue; a barrier
    (stream-ordered) tells ``dst`` that the field is complete.  CUDA tensors in, CUDA tensor out.
    ``compute(cov, z1, z2, local_axes, matrix)`` replaces the kernels (CPU tests with gloo; ``mode="nccl"``).
    ``reserve_sms`` (``mode="nccl"``, more than one piece): the persistent contraction normally holds EVERY SM with
    213 KB of shared memory per CTA, so NCCL's send/recv kernels could not start before it ends and nothing would
    overlap; while pieces are in flight the contraction runs on ``sm_count - reserve_sms`` CTAs instead.
    """
    import torch

    dist = _dist()
    rank, world = _rank_world(group)
    axes = list(axes)
    n0 = len(axes[0])
    dim = len(axes)
    fn = compute or (backend.summate_incompr_structured if incompr else backend.summate_structured)
    if world == 1:
        return fn(cov_samples, z_1, z_2, axes, matrix)
    rest = tuple(len(a) for a in axes[1:])
    lead = (dim,) if incompr else ()
    ranges = [shard_range(n0, r, world) for r in range(world)]
    lo, hi = ranges[rank]

    if mode == "p2p":
        if dst is None or compute is not None:
            raise ValueError('mode="p2p" gathers onto one rank and runs the CUDA kernels only')
        full, ptr = open_peer_field(lead + (n0,) + rest, dst, group)
        # Stream-ordered rendezvous (a one-element all-reduce; the host is not blocked): nobody stores into the
        # (reused) buffer before its previous consumer on `dst` is done, and `dst` continues only after every
        # rank's kernel -- enqueued before that rank's contribution -- has finished
        token = _token(torch.cuda.current_device())
        dist.all_reduce(token, group=group)
        if hi > lo:
            backend.structured_slab_into(cov_samples, z_1, z_2, axes, matrix, lo, hi, ptr, incompr=incompr)
        dist.all_reduce(token, group=group)
        return full

    if mode != "nccl":
        raise ValueError('mode must be "nccl" or "p2p"')
    receivers = list(range(world)) if dst is None else [dst]
    receiver = rank in receivers
    on_gpu = any(backend._is_cuda_tensor(x) for x in (cov_samples, z_1, z_2, *axes))
    cuts = {r: [a + ranges[r][0] for a in piece_bounds(ranges[r][1] - ranges[r][0], pieces)] for r in range(world)}
    n_pieces = max(len(c) - 1 for c in cuts.values())
    full = None
    keep, reqs = [], []
    main = side = None
    if on_gpu:
        dev = next(x.device for x in (cov_samples, z_1, z_2, *axes) if backend._is_cuda_tensor(x))
        main = torch.cuda.current_stream(dev)
        side = _side_stream(dev)
        side.wait_stream(main)
    shrink = on_gpu and compute is None and n_pieces > 1 and reserve_sms > 0
    if shrink:
        from . import _lib

        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        _lib.set_option("sk_grid", max(1, sms - int(reserve_sms)))
    for k in range(n_pieces):
        if shrink and k == n_pieces - 1:
            _lib.set_option("sk_grid", 0)      # nothing travels behind the last piece: all SMs again
        mine = None
        if k + 1 < len(cuts[rank]) and cuts[rank][k + 1] > cuts[rank][k]:
            a, b = cuts[rank][k], cuts[rank][k + 1]
            mine = fn(cov_samples, z_1, z_2, [axes[0][a:b]] + axes[1:], matrix)
            if isinstance(mine, np.ndarray):
                mine = torch.from_numpy(np.ascontiguousarray(mine))
            mine = mine.reshape(lead + (b - a,) + rest)
            keep.append(mine)
        if receiver and full is None:
            like = mine if mine is not None else torch.empty(0, dtype=torch.float64)
            full = torch.empty(lead + (n0,) + rest, dtype=torch.float64, device=like.device)
        if receiver and mine is not None:
            full[(slice(None),) * len(lead) + (slice(cuts[rank][k], cuts[rank][k + 1]),)].copy_(mine)
        ops = []
        # (several messages between one pair of ranks -- the components of a vector field -- keep their order)
        comps = [(c,) for c in range(dim)] if incompr else [()]
        if receiver:
            for r in range(world):
                if r != rank and k + 1 < len(cuts[r]) and cuts[r][k + 1] > cuts[r][k]:
                    for c in comps:
                        ops.append(dist.P2POp(dist.irecv, full[c + (slice(cuts[r][k], cuts[r][k + 1]),)],
                                              _global_rank(r, group), group))
        if mine is not None:
            for q in receivers:
                if q != rank:
                    for c in comps:
                        ops.append(dist.P2POp(dist.isend, mine[c], _global_rank(q, group), group))
        if not ops:
            continue
        if on_gpu:
            side.wait_stream(main)                     # the piece is complete on the compute stream
            with torch.cuda.stream(side):
                reqs += dist.batch_isend_irecv(ops)    # travels while the next piece contracts on `main`
        else:
            reqs += dist.batch_isend_irecv(ops)
    if shrink:
        _lib.set_option("sk_grid", 0)
    for req in reqs:
        req.wait()
    if on_gpu:
        main.wait_stream(side)
    del keep
    if full is None:
        return None
    return full


_SIDE = {}
_TOKEN = {}


def _token(index):
    import torch

    if index not in _TOKEN:
        _TOKEN[index] = torch.zeros(1, dtype=torch.float32, device=torch.device("cuda", index))
    return _TOKEN[index]


def _side_stream(dev):
    import torch

    key = (dev.type, dev.index)
    if key not in _SIDE:
        _SIDE[key] = torch.cuda.Stream(device=dev)
    return _SIDE[key]
