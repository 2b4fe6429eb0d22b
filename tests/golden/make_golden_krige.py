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
    meta = dict(name=name, kind="krige", cite=cite, var=float(krige.model.var), sill=float(krige.model.sill),
                cond_val=COND_VAL.tolist(), node_index=[int(i) for i in flat], exact=bool(krige.exact))
    if krige.exact and np.all(np.asarray(krige.cond_err) == 0):
        for got, val in zip(at_nodes, COND_VAL):        # the reference's own assertion (places=2)
            assert round(got - val, 2) == 0, (name, got, val)
    np.savez_compressed(os.path.join(HERE, "krige", name + ".npz"), meta=json.dumps(meta),
                        krig_mat=rec["krig_mat"], krig_vecs=rec["krig_vecs"], cond=rec["cond"],
                        field=rec["field"], error=rec["error"])
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
    m = gs.Exponential(dim=2, var=2, len_scale=10, anis=[0.9, 0.8], angles=[2, 1, 0.5])
    run("universal_linear_exponential_2d", gs.krige.Universal(m, COND_POS[:2], COND_VAL, "linear"), 2,
        "tests/test_krige.py:109-133 (test_universal)")


if __name__ == "__main__":
    main()
