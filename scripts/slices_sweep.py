"""GPU diagnostic: accuracy of the tensor-core walk half-move vs an f64 integration of the same draws, as a function of the
number of split-K slices of the projection GEMM (hx_ctx_set_tuning(1, slices))."""
import sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "h-emcee_b200")]
import hemcee_b200 as hx
from hemcee_b200 import _lib, _rt
from bench import make_cov
d = 1000; n = int(os.environ.get("N", 8192)); L = 10; eps = 0.1
dev = torch.device("cuda", 0)
lib = _lib.load(); ctx = _rt.ctx(dev)
_, prec = make_cov(d, 1e4)
tgt = hx.targets.GaussianFull(precision=prec)
k_init, k_run = hx.random.split(hx.random.PRNGKey(0), 2)
x0 = hx.random.normal(k_init, (2 * n, d), torch.float32)
keys = hx.random.split(k_run, 8)
s = hx.HamiltonianEnsembleSampler(2 * n, d, tgt, step_size=eps, L=L, adapt_inital_step_size=False, adapt_step_size=False, adapt_length=False, verbose=False)
s.run_mcmc(keys[5], x0, 30, warmup=0)
X = s.last_state
G1, G2 = X[:n].contiguous(), X[n:].contiguous()
LP1 = tgt.log_prob(G1)
P0 = hx.random.normal(keys[0], (n, n), torch.float32)
X2d = G2.double(); C = (X2d - X2d.mean(0)) / np.sqrt(n); S0 = P0.double() @ C; del P0
hx.config.update("enable_x64", True); hx.config.update("precision", "exact")
tg64 = hx.targets.GaussianFull(precision=prec)
q = G1.double().clone(); S = S0.clone(); dK = torch.empty(n, dtype=torch.float64, device=dev)
_lib.check(lib.hx_leapfrog_walk(ctx, 1, tg64.descriptor(torch.float64, dev), _rt.ptr(q), _rt.ptr(S), _rt.ptr(C.contiguous()), n, n, eps, L, _rt.ptr(dK), _rt.stream(dev)))
lp0 = tg64.log_prob(G1.double()); lpL = tg64.log_prob(q)
dH = (lp0 - lpL) + dK; la_t = torch.minimum(torch.zeros_like(dH), -dH)
hx.config.update("enable_x64", False)
sc = q.abs().max().item()
def rep(name, a):
    qq, la, pp, lp = [t.double() for t in a]
    print(f"{name:28s} |dq|/scale {(qq-q).abs().max().item()/sc:.2e} |dS| rel {(pp-S).abs().max().item()/S.abs().max().item():.2e} dlp rms {(lp-lpL).pow(2).mean().sqrt().item():.3e} "
          f"dla rms {(la-la_t).pow(2).mean().sqrt().item():.3e} mean {(la-la_t).mean().item():+.3e}", flush=True)
hx.config.update("precision", "exact")
rep("exact fp32 SIMT", hx.hmc_walk_move(G1, G2, eps, keys[0], tgt, tgt, L, LP1))
for form in ("stepwise", "propagator"):
    hx.config.update("leapfrog_form", form)
    for mode in ("bf16x3", "tf32x3"):
        hx.config.update("precision", mode)
        for g32 in (0, 1):
            lib.hx_ctx_set_tuning(ctx, 2, g32)
            rep(f"{mode} {form} gram={'tf32' if g32 else 'bf16'}", hx.hmc_walk_move(G1, G2, eps, keys[0], tgt, tgt, L, LP1))
