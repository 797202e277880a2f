"""world_size-2 `gloo` test (CPU) of the sampler's row-sharded orchestration: walkers of both
halves are split by rows over the ranks, each half-move updates the local rows against the all-gathered
complement, and — because the RNG is counter-based on GLOBAL row indices — the chain must equal the
unsharded oracle run.  The device kernels are replaced by a CPU stand-in (tests/_cpu_engine.py)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, move_kind, warm, out_dir):
    sys.path[:0] = [os.path.dirname(HERE), os.path.join(os.path.dirname(HERE), "h-emcee_b200"), HERE]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import hemcee_b200 as hx
    from hemcee_b200 import sampler as S
    import _cpu_engine
    from oracle import jaxrng as R, targets as OT
    hx.config.update("enable_x64", True)
    d, W, T = 3, 16, 6
    cov = OT.make_covariance_skewed(R.PRNGKey(7), d, 50.0, np.float64)
    otgt = OT.make_gaussian(cov, np.zeros(d))
    log_prob, make_move = _cpu_engine.install(hx, otgt)
    tgt = hx.targets.GaussianFull(cov=cov)
    tgt.log_prob = log_prob
    move = make_move(move_kind)
    s = S.HamiltonianEnsembleSampler(W, d, tgt, step_size=0.05, L=4, move=move, adapt_inital_step_size=False,
                                     adapt_step_size=bool(warm), adapt_length=False, group=dist.group.WORLD,
                                     exchange="nccl", verbose=False)   # collective path; 'peer' needs NVLink GPUs
    k_init, k_run = R.split(R.PRNGKey(0), 2)
    x0 = R.normal(k_init, (W, d), np.float64)
    chain = s.run_mcmc(k_run, x0, T, warmup=warm)
    np.save(os.path.join(out_dir, f"chain_{rank}.npy"), chain)
    np.save(os.path.join(out_dir, f"acc_{rank}.npy"), s.diagnostics_main["accepts"])
    dist.destroy_process_group()


@pytest.mark.parametrize("move_kind", ["walk", "side"])
@pytest.mark.parametrize("warm", [0, 5])
def test_row_sharded_sampler_matches_unsharded_oracle(tmp_path, move_kind, warm):
    from oracle import jaxrng as R, moves as OM, sampler as OS, targets as OT
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), move_kind, warm, str(tmp_path)), nprocs=world, join=True)
    d, W, T = 3, 16, 6
    cov = OT.make_covariance_skewed(R.PRNGKey(7), d, 50.0, np.float64)
    om = OM.hmc_walk_move if move_kind == "walk" else OM.hmc_side_move
    o = OS.HamiltonianEnsembleSampler(W, d, OT.make_gaussian(cov, np.zeros(d)), step_size=0.05, L=4, move=om,
                                      adapt_inital_step_size=False, adapt_step_size=bool(warm), adapt_length=False,
                                      dtype=np.float64)
    k_init, k_run = R.split(R.PRNGKey(0), 2)
    ref = o.run_mcmc(k_run, R.normal(k_init, (W, d), np.float64), T, warmup=warm)
    chains = [np.load(tmp_path / f"chain_{r}.npy") for r in range(world)]
    np.testing.assert_array_equal(chains[0], chains[1])                 # every rank assembles the same chain
    np.testing.assert_allclose(chains[0], ref, rtol=1e-10, atol=1e-12)
    np.testing.assert_array_equal(np.load(tmp_path / "acc_0.npy"), o.diagnostics_main["accepts"])
