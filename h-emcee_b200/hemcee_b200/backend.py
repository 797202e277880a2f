"""In-memory chain store with the reference Backend's interface
(reference src/hemcee/backend/backend.py:13-215): the chain lives on the HOST as a list of per-batch
arrays; `get_chain(discard, thin, flat)` slices `v[discard::thin]` (backend.py:63-66)."""
from __future__ import annotations

import numpy as np

from . import autocorr

__all__ = ["Backend", "FileBackend"]


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
        # which iterations the stored rows are: row k = iteration stored_first + k * stored_stride.  (0, 1) = every iteration
        # (the reference's store); the sampler's store="thinned" keeps only the iterations run_mcmc returns
        # (stored_first = warmup, stored_stride = thin_by) while `iteration` still counts the iterations that were RUN.
        self.stored_first, self.stored_stride = 0, 1
        self._last_sample = None
        self.stats = None               # on-device accumulators of the main phase (hemcee_b200.stats.ChainStats), if the sampler keeps them
        self.initialized = True

    def set_stored_layout(self, first: int, stride: int, last_sample=None):
        """Declare that the stored rows are iterations first, first + stride, ... (thinned store); `last_sample` =
        (coords[nwalkers, ndim], log_prob[nwalkers]) of the final iteration if that iteration itself is not stored."""
        self.stored_first, self.stored_stride = int(first), max(1, int(stride))
        self._last_sample = last_sample

    def _rows(self, discard, thin):
        """Translate the reference's `v[discard::thin]` over ALL iterations into a slice of the stored rows."""
        f, k = self.stored_first, self.stored_stride
        if f == 0 and k == 1:
            return slice(discard, None, thin)
        if discard < f:
            raise ValueError(f"iterations before {f} were not stored (thinned store); ask for discard >= {f}")
        if thin % k != 0:
            raise ValueError(f"only every {k}-th iteration was stored (thinned store): `thin` must be a multiple of {k}, got {thin}")
        if (discard - f) % k != 0:
            raise ValueError(f"iteration {discard} was not stored (thinned store keeps iterations {f} + j * {k})")
        return slice((discard - f) // k, None, thin // k)

    def get_value(self, name, flat=False, thin=1, discard=0):
        if self.iteration <= 0:
            raise AttributeError("you must run the sampler with 'store == True' before accessing the results")
        v = getattr(self, name)
        if name in ("chain", "log_prob"):
            if len(v) == 0:
                v = np.zeros((0, self.nwalkers, self.ndim) if name == "chain" else (0, self.nwalkers), dtype=np.float32)
            else:
                v = v[0] if len(v) == 1 else np.concatenate(v, axis=0)   # single slice: no copy
            v = v[self._rows(discard, thin)]
        else:
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
        last_stored = self.stored_first + (sum(len(c) for c in self.chain) - 1) * self.stored_stride
        if last_stored != it - 1:                            # thinned store whose final iteration is not a stored row
            if self._last_sample is None:
                raise AttributeError("the final iteration was not stored (thinned store) and no last sample was recorded")
            return self._last_sample
        c = self.chain[-1] if self.chain else None
        return c[-1], self.log_prob[-1][-1]

    def get_autocorr_time(self, discard=0, thin=1, **kwargs):
        if self.stats is not None and sum(len(c) for c in self.chain) == 0:
            # no chain was stored (store="none"): the estimate comes from the on-device accumulators of the main phase
            if thin != 1:
                raise ValueError("the on-device accumulators see every main-phase iteration: thin must be 1")
            kwargs.pop("has_walkers", None)
            return self.stats.integrated_time(**kwargs)
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


class FileBackend(Backend):
    """Backend that can be written to / restored from ONE file with the dataset and attribute names of the reference's
    HDFBackend group (reference src/hemcee/backend/hdf.py:99-128: datasets `chain[T, W, d]`, `log_prob[T, W]`, `accepted[W]`;
    attributes `version`, `nwalkers`, `ndim`, `iteration`, `iteration_warmup`, `iteration_main`, `warmup_end_index` with -1 for
    "warm-up not complete").  h5py is not a dependency here, so the container is NumPy's .npz; a maintainer with h5py can copy
    the arrays into an HDF5 group one to one.  During sampling it behaves like the in-memory Backend (the chain batches arrive
    from pinned host buffers); `save()` writes what has been stored so far.

    Differences from the reference's HDFBackend, stated rather than hidden: (1) the reference appends to the HDF5 file on every
    `save_slice` (hdf.py:155-170), so a crashed run keeps its completed batches; here that needs `autosave=True`, which REWRITES the
    .npz after every slice (atomic rename; cost grows with the stored chain -- meant for long batches, the sampler stores one slice
    per phase on a single GPU); (2) the file is .npz, which the reference's HDFBackend cannot open."""

    VERSION = 0

    def __init__(self, filename: str, name: str = "mcmc", read_only: bool = False, dtype=None, autosave: bool = False):
        super().__init__(dtype)
        self.filename, self.name, self.read_only, self.autosave = filename, name, read_only, autosave

    def save_slice(self, coords, log_prob, accepted, index: int):
        super().save_slice(coords, log_prob, accepted, index)
        if self.autosave and not self.read_only:
            self.save()

    def mark_warmup_end(self):
        super().mark_warmup_end()
        if self.autosave and not self.read_only and self.iteration > 0:
            self.save()

    def save(self):
        if self.read_only:
            raise RuntimeError("The backend has been loaded in read-only mode. Set `read_only = False` to make changes.")
        if not self.initialized:
            raise AttributeError("nothing to save: the backend has not been reset()")
        n = self.name
        empty_c = np.zeros((0, self.nwalkers, self.ndim), dtype=np.float32)
        chain = self.get_value("chain") if self.iteration > 0 and len(self.chain) else empty_c
        lp = self.get_value("log_prob") if self.iteration > 0 and len(self.log_prob) else empty_c[:, :, 0]
        attrs = dict(version=self.VERSION, nwalkers=self.nwalkers, ndim=self.ndim, iteration=self.iteration,
                     iteration_warmup=self.iteration_warmup, iteration_main=self.iteration_main,
                     warmup_end_index=-1 if self._warmup_end_index is None else self._warmup_end_index)
        arrays = {f"{n}/chain": np.asarray(chain), f"{n}/log_prob": np.asarray(lp), f"{n}/accepted": self.accepted}
        arrays.update({f"{n}/attrs/{k}": np.asarray(v) for k, v in attrs.items()})
        tmp = self.filename + ".tmp"
        with open(tmp, "wb") as f:
            np.savez(f, **arrays)
        import os
        os.replace(tmp, self.filename)                    # atomic: a crash during save keeps the previous file

    @classmethod
    def load(cls, filename: str, name: str = "mcmc", read_only: bool = True):
        b = cls(filename, name, read_only)
        with np.load(filename) as z:
            if f"{name}/chain" not in z:
                raise KeyError(f"no group {name!r} in {filename}")
            a = {k.split("/")[-1]: z[k].item() for k in z.files if k.startswith(f"{name}/attrs/")}
            b.reset(a["nwalkers"], a["ndim"])
            b.dtype = z[f"{name}/accepted"].dtype
            b.accepted = z[f"{name}/accepted"].copy()
            b.chain, b.log_prob = [z[f"{name}/chain"].copy()], [z[f"{name}/log_prob"].copy()]
            b.iteration, b.iteration_warmup, b.iteration_main = a["iteration"], a["iteration_warmup"], a["iteration_main"]
            b._warmup_end_index = None if a["warmup_end_index"] < 0 else a["warmup_end_index"]
        return b
