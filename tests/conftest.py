"""pytest configuration: markers, import paths, shared fixtures."""
import glob
import json
import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN_DIR = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _cuda_available():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _cuda_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    d = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    meta = json.loads(str(d["meta"]))
    return meta, {k: d[k] for k in d.files if k != "meta"}


def golden_names():
    names = sorted(os.path.splitext(os.path.basename(f))[0]
                   for f in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
    return [n for n in names if n != "config_modes"]  # mode sets only, no boundary record


@pytest.fixture(scope="session")
def gsb():
    """The product package with its CUDA library loaded (fails loudly if the .so is missing)."""
    import gstools_b200

    gstools_b200._lib.load()
    return gstools_b200


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle

    oracle.build()
    return oracle


def synth_modes(dim, n_modes, seed, len_scale=10.0):
    """Synthetic mode set shaped like RandMeth's: unit directions x heavy-tailed radii."""
    rs = np.random.RandomState(seed)
    z1, z2 = rs.normal(size=n_modes), rs.normal(size=n_modes)
    v = rs.normal(size=(dim, n_modes))
    v /= np.linalg.norm(v, axis=0)
    rad = np.minimum(np.abs(rs.standard_cauchy(n_modes)), 50.0) / len_scale
    return rad * v, z1, z2
