import sys, os, time
sys.path[:0] = ["/root/repo", "/root/repo/h-emcee_b200"]
import numpy as np, torch
import hemcee_b200 as hx
from oracle import jaxrng as R
keys = R.split(R.PRNGKey(0), 2)
x0 = R.normal(keys[0], (100, 2), np.float32)
for wu in (5000, 10000, 20000, 50000):
    t = time.time()
    s = hx.HamiltonianEnsembleSampler(100, 2, hx.targets.Rosenbrock(2), step_size=0.01, verbose=False)
    out = s.run_mcmc(keys[1], x0, 400, warmup=wu, thin_by=5)
    print(wu, float(s.found[0]), int(s.found[1]), float(s.found[0]) / 0.010159287601709366 - 1, round(time.time() - t, 1), flush=True)
