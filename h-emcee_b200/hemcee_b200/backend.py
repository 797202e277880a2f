"""In-memory chain store with the reference Backend's interface
(reference src/hemcee/backend/backend.py:13-215): the chain lives on the HOST as a list of per-batch
arrays; `get_chain(discard, thin, flat)` slices `v[discard::thin]` (backend.py:63-66)."""
from __future__ import annotations

import numpy as np

from . import autocorr

__all__ = ["Backend"]


class Backend:
    def __init__(self, dtype=None):
        self.dtype = np.float64 if dtype is None else dtype   # only the `accepted` counter (backend.py:28-30,47)
        self.initialized = False

    def reset(self, nwalkers: int, ndim: int):
        self.nwalkers, self.ndim = int(nwalkers), int(ndim)
        self.iteration = self.iteration_warmup = self.iteration_main = 0
        self._warmup_end_index = None
        self.accepted = np.zeros(self.nwalkers, dtype=self.dtype)
        self.chain, self.log_prob = [], []
        self.initialized = True

    def get_value(self, name, flat=False, thin=1, discard=0):
        if self.iteration <= 0:
            raise AttributeError("you must run the sampler with 'store == True' before accessing the results")
        v = getattr(self, name)
        if name in ("chain", "log_prob"):
            v = v[0] if len(v) == 1 else np.concatenate(v, axis=0)   # single slice: no copy
        v = v[discard::thin]
        if flat:
            s = list(v.shape[1:])
            s[0] = int(np.prod(v.shape[:2]))
            return v.reshape(s)
        return v

    def get_chain(self, **kwargs):
        return self.get_value("chain", **kwargs)

    def get_log_prob(self, **kwargs):
        return self.get_value("log_prob", **kwargs)

    def get_last_sample(self):
        if (not self.initialized) or self.iteration <= 0:
            raise AttributeError("you must run the sampler with 'store == True' before accessing the results")
        it = self.iteration
        return self.get_chain(discard=it - 1)[0], self.get_log_prob(discard=it - 1)[0]

    def get_autocorr_time(self, discard=0, thin=1, **kwargs):
        return thin * autocorr.integrated_time(self.get_chain(discard=discard, thin=thin), **kwargs)

    @property
    def shape(self):
        return self.nwalkers, self.ndim

    def save_slice(self, coords, log_prob, accepted, index: int):
        """coords[nsteps, nwalkers, ndim], log_prob[nsteps, nwalkers], accepted[nwalkers] (backend.py:174-211)."""
        coords = np.asarray(coords)
        nsteps = coords.shape[0]
        self.chain.append(coords)
        self.log_prob.append(np.asarray(log_prob))
        self.accepted = self.accepted + np.asarray(accepted, dtype=self.dtype)
        self.iteration = index + nsteps
        if self._warmup_end_index is None:
            self.iteration_warmup, self.iteration_main = self.iteration, 0
        else:
            self.iteration_warmup = min(self.iteration, self._warmup_end_index)
            self.iteration_main = max(0, self.iteration - self._warmup_end_index)

    def mark_warmup_end(self):
        self._warmup_end_index = self.iteration

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        pass
