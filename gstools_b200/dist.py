"""Multi-GPU sharding of the summation: one process per GPU, no traffic during the sum.

Every output element depends only on its own position and on the (tiny, replicated) mode set,
so the path shards into independent units:

* flat point sets       -> contiguous point ranges (:func:`shard_range`),
* structured meshes     -> slabs along axis 0, so each rank's output is one contiguous block of
                           the C-ordered field (reference layout: tools/geometric.py:340-356),
* ensembles of seeds    -> contiguous ranges of realisations (README.md:255-257 idiom).

NCCL (``torch.distributed``) is used only when the caller asks for the field as ONE array:
:func:`gather_field` all-gathers or gathers the per-rank slabs.  The compute functions are
injectable so the partition + gather logic is testable on CPU with the gloo backend.
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
    field (all-gather); ``dst=r`` -> only rank r does, others get ``None``.
    """
    import torch

    dist = _dist()
    rank, world = _rank_world(group)
    was_numpy = isinstance(local, np.ndarray)
    t = torch.from_numpy(np.ascontiguousarray(local)) if was_numpy else local
    if world == 1:
        return local
    t = t.movedim(axis, 0).contiguous()
    sizes = [hi - lo for lo, hi in (shard_range(n_total, r, world) for r in range(world))]
    if t.shape[0] != sizes[rank]:
        raise ValueError("local block does not match shard_range(n_total, rank, world)")
    # collectives want equal shapes on every rank: pad the short slabs by one row, trim after
    m = max(sizes)
    if t.shape[0] < m:
        pad = torch.zeros((m - t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        t = torch.cat([t, pad], dim=0)
    full = None
    if dst is None:
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t, group=group)
    else:
        parts = [torch.empty_like(t) for _ in range(world)] if rank == dst else None
        dist.gather(t, parts, dst=dst, group=group)
    if parts is not None:
        full = torch.cat([p[:k] for p, k in zip(parts, sizes)], dim=0).movedim(0, axis)
    if full is not None and was_numpy:
        return full.numpy()
    return full
