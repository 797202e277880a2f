"""hemcee_b200 — B200-native (sm_100a) drop-in for h-emcee's ensemble-Hamiltonian hot path.

Same emcee-style API as the reference package `hemcee` (reference src/hemcee/__init__.py:3-10) for
the path this project covers: `HamiltonianEnsembleSampler` with the Hamiltonian walk / side moves,
accept/reject, step-size and trajectory-length tuning hooks, chain getters and autocorrelation.
All numerical work is done by hand-written CUDA kernels in libhx.so through a ctypes C-ABI
(include/hx.h); there is no CPU fallback.
"""
from . import autocorr, random, targets
from ._lib import HxError
from ._rt import config
from .adaptation import ChEESParameters, DAParameters
from .backend import Backend, FileBackend
from .moves import hmc_side_move, hmc_walk_move
from .sampler import HamiltonianEnsembleSampler
from .ensemble import EnsembleSampler, side_move, stretch_move, walk_move

__version__ = "0.1.0"

__all__ = ["HamiltonianEnsembleSampler", "EnsembleSampler", "stretch_move", "side_move", "walk_move", "hmc_walk_move", "hmc_side_move", "DAParameters", "ChEESParameters",
           "Backend", "FileBackend", "autocorr", "random", "targets", "config", "HxError"]
