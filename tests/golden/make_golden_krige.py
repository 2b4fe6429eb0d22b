#!/usr/bin/env python
"""Generate tests/golden/krige/*.npz by running the UNMODIFIED reference in the build container.

    python tests/golden/make_golden_krige.py

The reference's kriging classes (src/gstools/krige/methods.py) are driven the way the reference's
tests/test_krige.py drives them (same conditioning data, models and grids, :22-110); the arrays
crossing ``_calc_field_krige_and_variance`` (krige/base.py:51-61) are recorded together with what
the native function returned (here: the CPU oracle through tools/refstubs/gstools_cython) and with
the known-answer facts those tests assert: the kriged field reproduces the conditioning values at
the conditioning nodes (places=2, test_krige.py:76-79, 104-107).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import refharness  # noqa: E402

gs = refharness.import_gstools()
from gstools.krige import base as kbase  # noqa: E402

import oracle  # noqa: E402

RECORD = []
_orig = kbase._calc_field_krige_and_variance


def _rec(krig_mat, krig_vecs, cond, num_threads=None):
    field, error = _orig(krig_mat, krig_vecs, cond, num_threads)
    RECORD.append(dict(krig_mat=np.array(krig_mat), krig_vecs=np.array(krig_vecs), cond=np.array(cond),
                       field=np.array(field), error=np.array(error)))
    return field, error


kbase._calc_field_krige_and_variance = _rec

DATA = np.array([[0.3, 1.2, 0.5, 0.47], [1.9, 0.6, 1.0, 0.56], [1.1, 3.2, 1.5, 0.74],
                 [3.3, 4.4, 2.0, 1.47], [4.7, 3.8, 2.5, 1.74]])      # test_krige.py:26-34
COND_POS = (DATA[:, 0], DATA[:, 1], DATA[:, 2])
COND_VAL = DATA[:, 3]
# coarser grids than test_krige.py:47-49 (every 10th / 5th node would miss the data nodes, so the
# data nodes are appended explicitly) to keep the fixtures small
AXES = [np.unique(np.concatenate([np.linspace(0, 5, 11), DATA[:, 0]])),
        np.unique(np.concatenate([np.linspace(0, 6, 7), DATA[:, 1]])),
        np.unique(np.concatenate([np.linspace(0, 7, 8), DATA[:, 2]]))]


def run(name, krige, dim, cite):
    pos = AXES[:dim]
    field, var = krige.structured(pos)
    rec = RECORD[-1]
    idx = tuple(np.searchsorted(AXES[t], DATA[:, t]) for t in range(dim))
    at_nodes = field[idx]
    flat = np.ravel_multi_index(idx, field.shape)
    model = krige.model
    spec = dict(kind=type(model).__name__, var=float(model.var), len_rescaled=float(model.len_rescaled),
                sill=float(model.sill), param=float(getattr(model, "alpha", 0.0)), exact=bool(krige.exact))
    meta = dict(name=name, kind="krige", cite=cite, var=float(model.var), sill=float(model.sill),
                cond_val=COND_VAL.tolist(), node_index=[int(i) for i in flat], exact=bool(krige.exact),
                spec=spec, unbiased=bool(krige.unbiased), shape=list(field.shape))
    # what the device-side right-hand-side generator needs, as the reference holds it
    iso_pos, _ = krige.pre_pos(pos, "structured")
    n_tail = krige.krige_size - krige.cond_no - int(krige.unbiased)
    tail = rec["krig_vecs"][krige.krige_size - n_tail:] if n_tail else np.zeros((0, iso_pos.shape[1]))
    from gstools.tools.geometric import matrix_isometrize
    extra = dict(cond_pos_iso=np.array(krige._krige_pos), pos_iso=np.array(iso_pos), tail_rows=np.array(tail),
                 matrix=matrix_isometrize(model.dim, model.angles, model.anis),
                 krige_var=np.array(var), **{f"axis{t}": np.array(a) for t, a in enumerate(pos)})
    # pin the numpy restatement of the covariance formulas on the reference's own right-hand sides
    kv_np = oracle.krige_vecs_np(spec["kind"], spec["var"], spec["len_rescaled"], spec["sill"],
                                 extra["cond_pos_iso"], extra["pos_iso"], krige.unbiased, tail,
                                 spec["param"], spec["exact"])
    assert np.max(np.abs(kv_np - rec["krig_vecs"])) <= 4e-16 * spec["sill"], (name, np.max(np.abs(kv_np - rec["krig_vecs"])))
    if krige.exact and np.all(np.asarray(krige.cond_err) == 0):
        for got, val in zip(at_nodes, COND_VAL):        # the reference's own assertion (places=2)
            assert round(got - val, 2) == 0, (name, got, val)
    np.savez_compressed(os.path.join(HERE, "krige", name + ".npz"), meta=json.dumps(meta),
                        krig_mat=rec["krig_mat"], krig_vecs=rec["krig_vecs"], cond=rec["cond"],
                        field=rec["field"], error=rec["error"], **extra)
    print(f"{name}: K={rec['krig_mat'].shape[0]} n={rec['krig_vecs'].shape[1]} "
          f"|M|max={np.abs(rec['krig_mat']).max():.3g}")


def main():
    for Model in (gs.Gaussian, gs.Exponential, gs.Spherical):
        for dim in (1, 2, 3):
            m = Model(dim=dim, var=2, len_scale=2, anis=[0.9, 0.8], angles=[2, 1, 0.5])
            run(f"simple_{Model.__name__.lower()}_{dim}d",
                gs.krige.Simple(m, COND_POS[:dim], COND_VAL, np.mean(COND_VAL)), dim,
                "tests/test_krige.py:57-79 (test_simple)")
            m = Model(dim=dim, var=5, len_scale=10, anis=[0.9, 0.8], angles=[2, 1, 0.5])
            run(f"ordinary_{Model.__name__.lower()}_{dim}d",
                gs.krige.Ordinary(m, COND_POS[:dim], COND_VAL), dim,
                "tests/test_krige.py:81-107 (test_ordinary)")
    for Model, kw in ((gs.Stable, dict(alpha=1.3)), (gs.Rational, dict(alpha=0.8)), (gs.Cubic, {}),
                      (gs.Linear, {}), (gs.Circular, {})):
        dim = 1 if Model is gs.Linear else 2
        m = Model(dim=dim, var=1.5, len_scale=4, nugget=0.1, anis=0.7, angles=0.4, **kw)
        run(f"ordinary_{Model.__name__.lower()}_{dim}d_exact",
            gs.krige.Ordinary(m, COND_POS[:dim], COND_VAL, exact=True), dim,
            "tests/test_krige.py:81-107 (test_ordinary), model list of covmodel/models.py")
    m = gs.Exponential(dim=2, var=2, len_scale=10, anis=[0.9, 0.8], angles=[2, 1, 0.5])
    run("universal_linear_exponential_2d", gs.krige.Universal(m, COND_POS[:2], COND_VAL, "linear"), 2,
        "tests/test_krige.py:109-133 (test_universal)")


if __name__ == "__main__":
    main()
