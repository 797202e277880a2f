"""GPU diagnostic (torchrun, N ranks): wall time of run_mcmc at c5 for batch sizes 1 and 10, and of its set-up steps."""
import os, sys, time
import numpy as np, torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "h-emcee_b200")]
import hemcee_b200 as hx
from bench import make_cov
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0))); torch.cuda.set_device(dev)
if world > 1: dist.init_process_group("nccl", device_id=dev)
W, d = 65536, 1000
tgt = hx.targets.GaussianFull(precision=make_cov(d, 1e4)[1])
k0, k1 = hx.random.split(hx.random.PRNGKey(0), 2)
x0 = hx.random.normal(k0, (W, d), torch.float32).cpu().pin_memory()
kw = dict(step_size=0.1, L=10, adapt_inital_step_size=False, adapt_step_size=False, adapt_length=False, verbose=False)
if world > 1: kw.update(group=dist.group.WORLD, chain_store="local", exchange="peer")
for bs in (1, 1, 10, 10, 1):
    s = hx.HamiltonianEnsembleSampler(W, d, tgt, **kw)
    torch.cuda.synchronize(dev)
    if world > 1: dist.barrier()
    t0 = time.perf_counter()
    out = s.run_mcmc(k1, x0, 10, warmup=0, batch_size=bs)
    torch.cuda.synchronize(dev)
    el = time.perf_counter() - t0
    if rank == 0: print(f"world {world} batch_size {bs}: {el * 1e3:.1f} ms for 10 iterations; timings {dict((k, round(v * 1e3, 1)) for k, v in s.timings.items())}", flush=True)
    del out, s
if world > 1: dist.destroy_process_group()
