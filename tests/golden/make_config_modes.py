#!/usr/bin/env python
"""Record the mode sets (cov_samples, z_1, z_2) the UNMODIFIED reference draws for
BASELINE.json configs 2-5, so tests and bench.py can run those configs on the GPU box where
/root/reference does not exist.

    python tests/golden/make_config_modes.py      # writes tests/golden/config_modes.npz

Radii of 3D / Matern models come from the reference's MCMC sampler (random/rng.py:38-104) run on
tools/refstubs/emcee, whose stream reproduces the reference's 16-digit 3D goldens.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import refharness  # noqa: E402

gs = refharness.import_gstools()
from gstools.field.generator import IncomprRandMeth, RandMeth  # noqa: E402
from gstools.random import MasterRNG  # noqa: E402

out = {}


def rec(tag, rm):
    out[tag + "_cov"] = np.array(rm._cov_sample)
    out[tag + "_z1"] = np.array(rm._z_1)
    out[tag + "_z2"] = np.array(rm._z_2)
    print(tag, rm._cov_sample.shape, "max|k| =", np.abs(rm._cov_sample).max())


# config 2: SRF Exponential 3D, mode_no=1000, seed 20170519
rec("c2", RandMeth(gs.Exponential(dim=3, var=1, len_scale=10), mode_no=1000, seed=20170519))
# config 3: Matern 2D (nu=1), mode_no=10000
rec("c3", RandMeth(gs.Matern(dim=2, var=1, len_scale=10, nu=1.0), mode_no=10000, seed=20170519))
# config 4: incompressible 3D Gaussian, mode_no=1000
rec("c4", IncomprRandMeth(gs.Gaussian(dim=3, var=1, len_scale=10), mode_no=1000, seed=20170519))
# config 5: ensemble seeds from MasterRNG(20170519) (README.md:255-257), first 8 realisations
seed = MasterRNG(20170519)
model = gs.Exponential(dim=3, var=1, len_scale=10)
seeds = []
for i in range(8):
    s = seed()
    seeds.append(s)
    rec(f"c5_{i}", RandMeth(model, mode_no=1000, seed=s))
out["c5_seeds"] = np.array(seeds)
np.savez_compressed(os.path.join(HERE, "config_modes.npz"), **out)
print("wrote config_modes.npz")
