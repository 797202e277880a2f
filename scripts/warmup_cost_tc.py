"""GPU diagnostic: cost per iteration of the adaptive warm-up (DA + ChEES) vs the main phase on the tensor-core path."""
import sys, os, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "h-emcee_b200")]
import hemcee_b200 as hx
from bench import make_cov
for W, d, iters in ((8192, 512, 20), (65536, 1000, 12)):
    t = hx.targets.GaussianFull(precision=make_cov(d, 1e4)[1])
    k0, k1 = hx.random.split(hx.random.PRNGKey(0), 2)
    x0 = hx.random.normal(k0, (W, d), torch.float32)
    for da, ch in ((False, False), (True, False), (True, True)):
        for rep in range(2):
            s = hx.HamiltonianEnsembleSampler(W, d, t, step_size=0.1, L=10, adapt_inital_step_size=False, adapt_step_size=da, adapt_length=ch,
                                              verbose=False, store="none")
            torch.cuda.synchronize(); t0 = time.perf_counter()
            s.run_mcmc(k1, x0, 0, warmup=iters)
            torch.cuda.synchronize(); el = time.perf_counter() - t0
        print(f"W={W} d={d} DA={da} ChEES={ch}: {el / iters * 1e3:.2f} ms per warm-up iteration, found={getattr(s, 'found', None)}", flush=True)
