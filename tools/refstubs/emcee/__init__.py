"""Minimal functional stand-in for ``emcee`` (test infrastructure).

Implements only what /root/reference/src/gstools/random/rng.py:77-104 uses: an
affine-invariant ensemble sampler with the Goodman & Weare stretch move (a = 2),
red/blue split of the walkers, ``run_mcmc`` / ``reset`` / ``get_chain``.  Written
from the published algorithm (Goodman & Weare 2010; Foreman-Mackey et al. 2013).
"""
import numpy as np

from .state import State

__version__ = "3.1.6+stub"
__all__ = ["EnsembleSampler", "State"]


class _StretchMove:
    def __init__(self, a=2.0, nsplits=2, randomize_split=True):
        self.a = a
        self.nsplits = nsplits
        self.randomize_split = randomize_split

    def get_proposal(self, s, c, random):
        c = np.concatenate(c, axis=0)
        ns, nc = len(s), len(c)
        ndim = s.shape[1]
        zz = ((self.a - 1.0) * random.rand(ns) + 1) ** 2.0 / self.a
        factors = (ndim - 1.0) * np.log(zz)
        rint = random.randint(nc, size=(ns,))
        return c[rint] - (c[rint] - s) * zz[:, None], factors

    def propose(self, sampler, state):
        random = sampler._random
        nwalkers, _ = state.coords.shape
        accepted = np.zeros(nwalkers, dtype=bool)
        all_inds = np.arange(nwalkers)
        inds = all_inds % self.nsplits
        if self.randomize_split:
            random.shuffle(inds)
        for split in range(self.nsplits):
            s1 = inds == split
            sets = [state.coords[inds == j] for j in range(self.nsplits)]
            s = sets[split]
            c = sets[:split] + sets[split + 1:]
            q, factors = self.get_proposal(s, c, random)
            new_log_probs = sampler.compute_log_prob(q)
            for j, f, nlp in zip(all_inds[s1], factors, new_log_probs):
                lnpdiff = f + nlp - state.log_prob[j]
                if lnpdiff > np.log(random.rand()):
                    accepted[j] = True
            m1 = s1 & accepted
            m2 = accepted[s1]
            state.coords[m1] = q[m2]
            state.log_prob[m1] = new_log_probs[m2]
        return state, accepted


class EnsembleSampler:
    def __init__(self, nwalkers, ndim, log_prob_fn, vectorize=False, **kwargs):
        self.nwalkers = nwalkers
        self.ndim = ndim
        self.log_prob_fn = log_prob_fn
        self.vectorize = vectorize
        self._moves = [_StretchMove()]
        self._weights = [1.0]
        self._random = np.random.mtrand.RandomState()
        self._chain = []

    @property
    def random_state(self):
        return self._random.get_state()

    @random_state.setter
    def random_state(self, state):
        try:
            self._random.set_state(state)
        except Exception:  # same silent behaviour as the real sampler
            pass

    def reset(self):
        self._chain = []

    def compute_log_prob(self, coords):
        p = coords
        if np.any(np.isinf(p)):
            raise ValueError("At least one parameter value was infinite")
        if np.any(np.isnan(p)):
            raise ValueError("At least one parameter value was NaN")
        if self.vectorize:
            results = self.log_prob_fn(p)
        else:
            results = [self.log_prob_fn(x) for x in p]
        try:
            log_prob = np.array([float(np.asarray(l).reshape(-1)[0]) for l in results])
        except (IndexError, TypeError):
            log_prob = np.array([float(l) for l in results])
        if np.any(np.isnan(log_prob)):
            raise ValueError("Probability function returned NaN")
        return log_prob

    def run_mcmc(self, initial_state, nsteps, **kwargs):
        state = State(initial_state, copy=True)
        if np.shape(state.coords) != (self.nwalkers, self.ndim):
            raise ValueError("incompatible input dimensions")
        self.random_state = state.random_state
        if state.log_prob is None:
            state.log_prob = self.compute_log_prob(state.coords)
        if np.shape(state.log_prob) != (self.nwalkers,):
            raise ValueError("incompatible input dimensions")
        if np.any(np.isnan(state.log_prob)):
            raise ValueError("The initial log_prob was NaN")
        for _ in range(int(nsteps)):
            move = self._random.choice(self._moves, p=self._weights)
            state, _accepted = move.propose(self, state)
            state.random_state = self.random_state
            self._chain.append(state.coords.copy())
        return state

    def get_chain(self, flat=False, **kwargs):
        chain = np.array(self._chain).reshape((-1, self.nwalkers, self.ndim))
        if flat:
            return chain.reshape((-1, self.ndim))
        return chain
