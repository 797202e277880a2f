"""Test-only CPU stand-in for libhx so the HOST-side orchestration of the product sampler (row sharding,
key order, all-gather per half-move, chain assembly) can be exercised with world_size-2 `gloo` on a box
without GPUs.  It routes the three device entry points through the oracle; the product never does this."""
import numpy as np
import torch

from oracle import jaxrng as R, moves as OM


def install(hx, oracle_target):
    """Monkeypatch hemcee_b200 so tensors live on the CPU and kernels are replaced by oracle calls."""
    from hemcee_b200 import _rt, sampler as S, random as hrandom, adaptation
    olp, ograd = oracle_target
    cpu = torch.device("cpu")
    _rt.device = lambda dev=None: cpu
    S._rt.device = _rt.device
    hrandom.split = lambda key, num=2, device=None, impl=None: R.split(np.asarray(key, np.uint32), num, impl or "partitionable")
    S.hrandom.split = hrandom.split
    adaptation.hrandom.split = hrandom.split

    def log_prob(X):
        return torch.from_numpy(olp(X.numpy()))

    def make_move(kind):
        def move(group1, group2, step_size, key, log_prob_, grad_log_prob, L, log_prob_group1, *, row0=0, n=None, impl=None):
            g1, g2 = group1.numpy(), group2.numpy()
            dt = g1.dtype.type
            n_ = g2.shape[0] if n is None else n
            rows = g1.shape[0]
            impl = impl or "partitionable"
            if kind == "walk":
                C = (g2 - np.mean(g2, axis=0, dtype=dt)) / np.sqrt(dt(n_))
                p0 = R.normal(key, (n_, n_), dt, impl)[row0:row0 + rows]
                q, p, pp = OM.leapfrog_walk_move(g1, p0, ograd, step_size, L, C)
                K0, K1 = dt(0.5) * np.sum(p0 ** 2, axis=1), dt(0.5) * np.sum(p ** 2, axis=1)
            else:
                keys = R.split(key, n_ + 2, impl)
                ch = OM.side_choices(keys[row0:row0 + rows], n_, impl)
                D = (g2[ch[:, 0]] - g2[ch[:, 1]]) / np.sqrt(dt(2 * g1.shape[1]))
                p0 = R.normal(keys[n_], (n_,), dt, impl)[row0:row0 + rows]
                q, p, pp = OM.leapfrog_side_move(g1, p0, ograd, step_size, L, D)
                K0, K1 = dt(0.5) * p0 ** 2, dt(0.5) * p ** 2
            lpq = olp(q)
            dH = (-lpq + K1) - (-log_prob_group1.numpy() + K0)
            dH = np.where(np.isnan(dH), dt(np.inf), dH)
            la = np.minimum(dt(0), -dH).astype(dt)
            return tuple(torch.from_numpy(np.ascontiguousarray(a)) for a in (q, la, pp, lpq))
        move.__name__ = "hmc_walk_move" if kind == "walk" else "hmc_side_move"
        return move

    def accept_proposal(current, proposed, lp_current, lp_proposed, log_accept_prob, key, accepts=None, *, row0=0,
                        n=None, impl=None):
        rows = current.shape[0]
        n_ = rows if n is None else n
        dt = current.numpy().dtype.type
        logu = OM.accept_uniform_log(key, n_, dt, impl or "partitionable")[row0:row0 + rows]
        mask = torch.from_numpy(logu < log_accept_prob.numpy())
        current[mask] = proposed[mask]
        lp_current[mask] = lp_proposed[mask]
        if accepts is not None:
            accepts += mask.to(torch.int32)
        return current, lp_current, accepts

    S.accept_proposal = accept_proposal
    S.walk_prefetch = lambda *a, **k: None      # device-side optimisation hint: nothing to do for the stand-in
    return log_prob, make_move
