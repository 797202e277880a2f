"""Multi-GPU (NCCL) run of the row-sharded sampler: needs >= 2 visible GPUs, skipped otherwise.
G = 2 ranks must reproduce the single-GPU chain bit-for-bit (shard invariance)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path[:0] = [{root!r}, os.path.join({root!r}, "h-emcee_b200")]
import hemcee_b200 as hx
from oracle import jaxrng as R, targets as OT
rank = int(os.environ["RANK"]); torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
d, W, T = 24, 256, 5
cov = OT.make_covariance_skewed(R.PRNGKey(7), d, 100.0, np.float64)
tgt = hx.targets.GaussianFull(cov=cov)
k_init, k_run = R.split(R.PRNGKey(0), 2)
x0 = R.normal(k_init, (W, d), np.float32)
for exchange in ("peer", "nccl"):
    for move in (hx.hmc_walk_move, hx.hmc_side_move):
        s = hx.HamiltonianEnsembleSampler(W, d, tgt, step_size=0.05, L=4, move=move, adapt_inital_step_size=False,
                                          adapt_step_size=False, adapt_length=False, group=dist.group.WORLD,
                                          exchange=exchange, verbose=False)
        chain = s.run_mcmc(k_run, x0, T, warmup=0, batch_size=2)
        if rank == 0:
            np.save(os.path.join({out!r}, "chain_" + exchange + "_" + move.__name__ + ".npy"), chain)
            np.save(os.path.join({out!r}, "acc_" + exchange + "_" + move.__name__ + ".npy"), s.diagnostics_main["accepts"])
# tensor-core walk path, sharded, peer exchange vs single GPU (rows = 128 per rank)
hx.config.update("precision", "bf16x3")
d2, W2 = 64, 512
cov2 = OT.make_covariance_skewed(R.PRNGKey(8), d2, 100.0, np.float64)
tgt2 = hx.targets.GaussianFull(cov=cov2)
x02 = R.normal(k_init, (W2, d2), np.float32)
s = hx.HamiltonianEnsembleSampler(W2, d2, tgt2, step_size=0.03, L=4, adapt_inital_step_size=False, adapt_step_size=False,
                                  adapt_length=False, group=dist.group.WORLD, exchange="peer", verbose=False)
chain = s.run_mcmc(k_run, x02, 3, warmup=0)
if rank == 0:
    np.save(os.path.join({out!r}, "chain_tc.npy"), chain)
dist.destroy_process_group()
'''


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_chain_equals_single_gpu(hx, tmp_path):
    from oracle import jaxrng as R, targets as OT
    script = tmp_path / "run.py"
    script.write_text(SCRIPT.format(root=ROOT, out=str(tmp_path)))
    with socket.socket() as sck:
        sck.bind(("127.0.0.1", 0)); port = sck.getsockname()[1]
    subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                    "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)], check=True, timeout=600)
    d, W, T = 24, 256, 5
    cov = OT.make_covariance_skewed(R.PRNGKey(7), d, 100.0, np.float64)
    tgt = hx.targets.GaussianFull(cov=cov)
    k_init, k_run = R.split(R.PRNGKey(0), 2)
    x0 = R.normal(k_init, (W, d), np.float32)
    for move in (hx.hmc_walk_move, hx.hmc_side_move):
        s = hx.HamiltonianEnsembleSampler(W, d, tgt, step_size=0.05, L=4, move=move, adapt_inital_step_size=False,
                                          adapt_step_size=False, adapt_length=False, verbose=False)
        ref = s.run_mcmc(k_run, x0, T, warmup=0)
        for exchange in ("peer", "nccl"):
            got = np.load(tmp_path / f"chain_{exchange}_{move.__name__}.npy")
            np.testing.assert_array_equal(got, ref)
            np.testing.assert_array_equal(np.load(tmp_path / f"acc_{exchange}_{move.__name__}.npy"), s.diagnostics_main["accepts"])
    # tensor-core path: the sharded chain equals the single-GPU tensor-core chain bit for bit
    hx.config.update("precision", "bf16x3")
    try:
        d2, W2 = 64, 512
        cov2 = OT.make_covariance_skewed(R.PRNGKey(8), d2, 100.0, np.float64)
        tgt2 = hx.targets.GaussianFull(cov=cov2)
        x02 = R.normal(k_init, (W2, d2), np.float32)
        s = hx.HamiltonianEnsembleSampler(W2, d2, tgt2, step_size=0.03, L=4, adapt_inital_step_size=False, adapt_step_size=False,
                                          adapt_length=False, verbose=False)
        ref = s.run_mcmc(k_run, x02, 3, warmup=0)
        np.testing.assert_array_equal(np.load(tmp_path / "chain_tc.npy"), ref)
    finally:
        hx.config.update("precision", "auto")
