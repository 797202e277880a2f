"""GPU diagnostic: cost per iteration of the adaptive warm-up (DA + ChEES, host-driven) vs the main phase."""
import sys, os, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "h-emcee_b200")]
import hemcee_b200 as hx
for name, tgt, W, d, eps in (("c3 rosenbrock d=50 W=8192", None, 8192, 50, 0.01), ("c2 gauss d=100 W=1024", "g", 1024, 100, 0.1), ("quickstart d=3 W=20", "iso", 20, 3, 0.1)):
    if tgt is None: t = hx.targets.Rosenbrock(d)
    elif tgt == "iso": t = hx.targets.GaussianIso(d)
    else:
        from bench import make_cov
        t = hx.targets.GaussianFull(precision=make_cov(d, 1e4)[1])
    k0, k1 = hx.random.split(hx.random.PRNGKey(0), 2)
    x0 = hx.random.normal(k0, (W, d), torch.float32)
    if tgt is None: x0 = x0 * 0.3 + 1.0
    for warm, ns in ((0, 200), (200, 200)):
        s = hx.HamiltonianEnsembleSampler(W, d, t, step_size=eps, L=10, adapt_inital_step_size=False, adapt_step_size=bool(warm), adapt_length=bool(warm), verbose=False)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        s.run_mcmc(k1, x0, ns, warmup=warm)
        torch.cuda.synchronize(); el = time.perf_counter() - t0
        print(f"{name}: warmup={warm} samples={ns}: {el*1e3:.1f} ms total, found={getattr(s,'found',None)}", flush=True)
