import sys; sys.path[:0] = ["/root/repo", "/root/repo/h-emcee_b200"]
import numpy as np, hemcee_b200 as hx
from bench import make_cov
d, W = 256, 4096
cov, _ = make_cov(d, 1e3)
tgt = hx.targets.GaussianFull(cov=cov)
keys = hx.random.split(hx.random.PRNGKey(0), 2)
x0 = hx.random.normal(keys[0], (W, d))
s = hx.HamiltonianEnsembleSampler(W, d, tgt, step_size=0.1, L=10, store="none", online_stats=True, verbose=False)
s.run_mcmc(keys[1], x0, num_samples=300, warmup=100)
print(s.found, s.stats.integrated_time(quiet=True), s.stats.mean()[:3], s.stats.var()[:3], np.diag(cov)[:3])
