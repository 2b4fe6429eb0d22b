"""``emcee.state.State`` stand-in (rng.py:15)."""
from copy import deepcopy

import numpy as np


class State:
    __slots__ = ("coords", "log_prob", "blobs", "random_state")

    def __init__(self, coords, log_prob=None, blobs=None, random_state=None, copy=False):
        dc = deepcopy if copy else (lambda x: x)
        if hasattr(coords, "coords"):
            self.coords = dc(coords.coords)
            self.log_prob = dc(coords.log_prob)
            self.blobs = dc(coords.blobs)
            self.random_state = dc(coords.random_state)
            return
        self.coords = dc(np.atleast_2d(coords))
        self.log_prob = dc(log_prob)
        self.blobs = dc(blobs)
        self.random_state = dc(random_state)
