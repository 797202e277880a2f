"""NumPy restatement of the reference's ensemble-HMC driver, chain store and MH accept.

TEST INFRASTRUCTURE (see oracle/__init__.py).

  HamiltonianEnsembleSampler.run_mcmc   reference src/hemcee/samplers/hamiltonian_ensemble.py:64-139
  warm-up scan body / key order         reference src/hemcee/samplers/hamiltonian_ensemble.py:165-248
  main scan body / key order            reference src/hemcee/samplers/hamiltonian_ensemble.py:284-343
  BaseSampler validation / getters      reference src/hemcee/samplers/sampler.py:24-102
  batched_scan / calculate_batch_size   reference src/hemcee/samplers/sampler_utils.py:10-88
  Backend                               reference src/hemcee/backend/backend.py:13-215

The per-iteration loop is a plain Python loop (the reference's `lax.scan`), so
only small cases are practical; that is all the parity tests need.
"""
from __future__ import annotations

import numpy as np

from . import adaptation, autocorr, jaxrng, moves


def calculate_batch_size(total_chains, dim, total_iterations, user_batch_size=None):
    """sampler_utils.py:10-26."""
    if user_batch_size is not None:
        return user_batch_size
    return min(max(50, 10000 // (total_chains * dim)), total_iterations)


class Backend:
    """backend.py:13-215 (in-memory chain store)."""

    def __init__(self, dtype=None):
        self.dtype = np.float64 if dtype is None else dtype

    def reset(self, nwalkers, ndim):
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
            v = np.concatenate(v, axis=0)
        v = v[discard::thin]
        if flat:
            s = list(v.shape[1:])
            s[0] = int(np.prod(v.shape[:2]))
            return v.reshape(s)
        return v

    def get_chain(self, **kw):
        return self.get_value("chain", **kw)

    def get_log_prob(self, **kw):
        return self.get_value("log_prob", **kw)

    def get_last_sample(self):
        it = self.iteration
        return self.get_chain(discard=it - 1)[0], self.get_log_prob(discard=it - 1)[0]

    def get_autocorr_time(self, discard=0, thin=1, **kw):
        return thin * autocorr.integrated_time(self.get_chain(discard=discard, thin=thin), **kw)

    @property
    def shape(self):
        return self.nwalkers, self.ndim

    def save_slice(self, coords, log_prob, accepted, index):
        nsteps = coords.shape[0]
        self.chain.append(np.asarray(coords))
        self.log_prob.append(np.asarray(log_prob))
        self.accepted = self.accepted + accepted
        self.iteration = index + nsteps
        if self._warmup_end_index is None:
            self.iteration_warmup, self.iteration_main = self.iteration, 0
        else:
            self.iteration_warmup = min(self.iteration, self._warmup_end_index)
            self.iteration_main = max(0, self.iteration - self._warmup_end_index)

    def mark_warmup_end(self):
        self._warmup_end_index = self.iteration


class HamiltonianEnsembleSampler:
    """hamiltonian_ensemble.py:12-346 with `log_prob` given as an oracle target pair
    `(log_prob[n,d]->[n], grad_log_prob[n,d]->[n,d])` instead of a JAX callable."""

    def __init__(self, total_chains, dim, log_prob, step_size=0.1, L=10, move=moves.hmc_walk_move,
                 backend=None, adapt_inital_step_size=True, adapt_step_size=True, adapt_length=True,
                 dtype=np.float32, impl=jaxrng.DEFAULT_IMPL, verbose=False):
        if total_chains < 4:
            raise ValueError("`total_chains` must be at least 4 to form meaningful ensemble groups")
        self.total_chains, self.dim = int(total_chains), int(dim)
        self.move = move
        self.backend = Backend() if backend is None else backend
        self.backend.reset(total_chains, dim)
        self.log_prob, self.grad_log_prob = log_prob
        self.dt = np.dtype(dtype).type
        self.impl = impl
        self.verbose = verbose
        self.step_size, self.L = step_size, L
        self.adapt_inital_step_size = adapt_inital_step_size
        self.adapter = adaptation.select_adapter(adapt_step_size, adapt_length, move, step_size, L, self.dt)
        self.adapter_state = self.adapter.init(self.dim)
        self.diagnostics_warmup = self.diagnostics_main = None
        self.found = None

    # -- sampler.py:61-102 ---------------------------------------------------------------
    def _validate(self, num_samples, warmup, thin_by, initial_state):
        if thin_by < 1:
            raise ValueError("`thin_by` must be 1 or greater.")
        if warmup < 0:
            raise ValueError("`warmup` must be 0 or greater.")
        if num_samples < 0:
            raise ValueError("`num_samples` must be 0 or greater.")
        if initial_state.shape != (self.total_chains, self.dim):
            raise ValueError("`inital_state` needs to have shape ")

    def get_chain(self, discard=0, thin=1, flat=False):
        return self.backend.get_chain(discard=discard, thin=thin, flat=flat)

    def get_logprob(self, discard=0, thin=1, flat=False):
        return self.backend.get_log_prob(discard=discard, thin=thin, flat=flat)

    def get_acceptance_prob(self):
        return self.backend.accepted / self.backend.iteration

    def get_autocorr(self, discard, thin):
        return autocorr.integrated_time(self.get_chain(discard=discard, thin=thin))

    # -- hamiltonian_ensemble.py:64-139 -----------------------------------------------------
    def run_mcmc(self, key, initial_state, num_samples, warmup=1000, thin_by=1, batch_size=None):
        dt = self.dt
        keys = jaxrng.split(key, 3, self.impl)                                      # :90
        initial_state = np.asarray(initial_state, dtype=dt)
        self._validate(num_samples, warmup, thin_by, initial_state)                 # :95
        n1 = self.total_chains // 2
        g1, g2 = initial_state[:n1], initial_state[n1:]                             # :107-108
        wb = calculate_batch_size(self.total_chains, self.dim, warmup, batch_size)
        mb = calculate_batch_size(self.total_chains, self.dim, num_samples * thin_by, batch_size)
        if self.adapt_inital_step_size:                                             # :115-118
            self.step_size = adaptation.find_reasonable_step_size(
                keys[0], g1, g2, self.log_prob, self.grad_log_prob, self.step_size, self.move, self.impl)
            if self.verbose:
                print(f"Using inital step size of {float(self.step_size)!r}")
        if warmup > 0:                                                              # :121-129
            g1, g2 = self._mcmc_warmup(keys[1], g1, g2, warmup, wb)
            step_size, L = self.adapter.finalize(self.adapter_state)
            if self.verbose:
                print(f"Found step size {float(step_size)!r} and integration length: {L}")
        else:
            step_size, L = self.step_size, self.L                                   # :131-132
        self.found = (step_size, L)
        self._mcmc_main(keys[2], g1, g2, num_samples, thin_by, step_size, L, mb, warmup)   # :136
        return self.get_chain(discard=warmup, thin=thin_by)                         # :139

    def _half(self, ga, gb, step, L, k_move, k_acc, lpa, adapter_state=None):
        prop, la, pproj, lpp = self.move(ga, gb, step, k_move, self.log_prob, self.grad_log_prob, L, lpa,
                                         impl=self.impl)
        if adapter_state is not None:                                               # :176-182 (before accept)
            adapter_state = self.adapter.update(
                state=adapter_state, log_accept_rate=la, position_current=ga,
                position_proposed=prop, momentum_proposed=pproj, group2=gb)
        new_g, acc = moves.accept_proposal(ga, prop, la, k_acc, self.impl)
        new_lp, _ = moves.accept_proposal(lpa, lpp, la, k_acc, self.impl)           # same key (:184-185)
        return new_g, new_lp, acc, adapter_state

    def _scan(self, body, carry, keys, batch_size, offset):
        """sampler_utils.py:29-88."""
        T = keys.shape[0]
        for start in range(0, T, batch_size):
            stop = min(start + batch_size, T)
            coords, lps, accs = [], [], []
            for t in range(start, stop):
                carry, (c, lp, a) = body(carry, keys[t])
                coords.append(c); lps.append(lp); accs.append(a)
            coords, lps = np.stack(coords), np.stack(lps)
            if np.any(np.isnan(lps)) or np.any(np.isinf(lps)):                      # :74-75
                raise ValueError("Log probability returned NaN or inf value, please check your log probability function")
            self.backend.save_slice(coords, lps, np.sum(np.stack(accs), axis=0), offset + start)
        return carry

    def _mcmc_warmup(self, key, g1, g2, warmup, batch_size):
        """hamiltonian_ensemble.py:141-248; key order move1, accept1, move2, accept2 = keys[0..3]."""
        def body(carry, k):
            g1, g2, lp1, lp2, st, accepts = carry
            step, L = self.adapter.value(st)                                        # :169
            g1, lp1, a1, st = self._half(g1, g2, step, L, k[0], k[1], lp1, st)
            step, L = self.adapter.value(st)                                        # :189
            g2, lp2, a2, st = self._half(g2, g1, step, L, k[2], k[3], lp2, st)
            a = np.concatenate([a1, a2])
            return (g1, g2, lp1, lp2, st, accepts + a), (np.concatenate([g1, g2]), np.concatenate([lp1, lp2]), a)

        lp1, lp2 = self.log_prob(g1), self.log_prob(g2)                             # :224-225
        keys = jaxrng.split(key, 4 * warmup, self.impl).reshape(warmup, 4, 2)       # :227-228
        carry = (g1, g2, lp1, lp2, self.adapter_state, np.zeros(self.total_chains))
        g1, g2, lp1, lp2, st, accepts = self._scan(body, carry, keys, batch_size, 0)
        self.diagnostics_warmup = {"accepts": accepts, "acceptance_rate": accepts / warmup}
        self.adapter_state = st
        self.backend.mark_warmup_end()
        return g1, g2

    def _mcmc_main(self, key, g1, g2, num_samples, thin_by, step_size, L, batch_size, warmup_offset):
        """hamiltonian_ensemble.py:250-343; key order move1=keys[0], accept1=keys[2], move2=keys[1], accept2=keys[3]."""
        def body(carry, k):
            g1, g2, lp1, lp2, accepts = carry
            g1, lp1, a1, _ = self._half(g1, g2, step_size, L, k[0], k[2], lp1)
            g2, lp2, a2, _ = self._half(g2, g1, step_size, L, k[1], k[3], lp2)
            a = np.concatenate([a1, a2])
            return (g1, g2, lp1, lp2, accepts + a), (np.concatenate([g1, g2]), np.concatenate([lp1, lp2]), a)

        lp1, lp2 = self.log_prob(g1), self.log_prob(g2)                             # :325-326
        total = num_samples * thin_by
        keys = jaxrng.split(key, 4 * total, self.impl).reshape(total, 4, 2)         # :328-330
        carry = (g1, g2, lp1, lp2, np.zeros(self.total_chains))
        if total == 0:
            self.diagnostics_main = {"accepts": carry[-1], "acceptance_rate": carry[-1]}
            return
        *_, accepts = self._scan(body, carry, keys, batch_size, warmup_offset)
        self.diagnostics_main = {"accepts": accepts, "acceptance_rate": accepts / total}


class EnsembleSampler:
    """reference src/hemcee/samplers/ensemble.py:9-140 (derivative-free affine-invariant ensemble sampler)."""

    def __init__(self, total_chains, dim, log_prob, move=moves.stretch_move, backend=None, dtype=np.float32,
                 impl=None):
        from . import jaxrng
        if total_chains < 4:
            raise ValueError("`total_chains` must be at least 4 to form meaningful ensemble groups")
        self.total_chains, self.dim = int(total_chains), int(dim)
        self.log_prob = log_prob[0] if isinstance(log_prob, tuple) else log_prob
        self.move = move
        self.dtype = dtype
        self.impl = impl or jaxrng.DEFAULT_IMPL
        self.backend = Backend() if backend is None else backend
        self.backend.reset(total_chains, dim)
        self.diagnostics_main = None

    def get_chain(self, discard=0, thin=1, flat=False):
        return self.backend.get_chain(discard=discard, thin=thin, flat=flat)

    def run_mcmc(self, key, initial_state, num_samples, warmup=1000, thin_by=1, batch_size=None, **kwargs):
        from . import jaxrng
        total = warmup + num_samples * thin_by                                  # :66
        x = np.asarray(initial_state, dtype=self.dtype)
        n1 = self.total_chains // 2
        keys = jaxrng.split(key, total * 2, self.impl).reshape(total, 2, 2)     # :79-81
        g1, g2 = x[:n1].copy(), x[n1:].copy()
        lp1, lp2 = self.log_prob(g1), self.log_prob(g2)
        accepts = np.zeros(self.total_chains)
        chain = np.empty((total, self.total_chains, self.dim), self.dtype)
        lps = np.empty((total, self.total_chains), self.dtype)
        for t in range(total):                                                  # body :97-122
            k = keys[t]
            prop, la, lpp = self.move(g1, g2, k[0], self.log_prob, lp1, impl=self.impl, **kwargs)
            g1, a1 = moves.accept_proposal(g1, prop, la, k[0], self.impl)
            lp1, _ = moves.accept_proposal(lp1, lpp, la, k[0], self.impl)
            prop, la, lpp = self.move(g2, g1, k[1], self.log_prob, lp2, impl=self.impl, **kwargs)
            g2, a2 = moves.accept_proposal(g2, prop, la, k[1], self.impl)
            lp2, _ = moves.accept_proposal(lp2, lpp, la, k[1], self.impl)
            accepts += np.concatenate([a1, a2])
            chain[t] = np.concatenate([g1, g2])
            lps[t] = np.concatenate([lp1, lp2])
            if not np.all(np.isfinite(lps[t])):
                raise ValueError("Log probability returned NaN or inf value, please check your log probability function")
        self.backend.save_slice(chain, lps, accepts, 0)
        self.diagnostics_main = {"accepts": accepts, "acceptance_rate": accepts / total if total else accepts}
        return self.get_chain(discard=warmup, thin=thin_by)
