"""GPU diagnostic: run-to-run determinism of hx_run_iterations at c5 scale, both leapfrog forms."""
import sys, os, ctypes as C
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "h-emcee_b200")]
import hemcee_b200 as hx
from hemcee_b200 import _lib, _rt
from bench import make_cov
d = 1000; W = int(os.environ.get("W", 65536)); n = W // 2; L = 10; eps = 0.1; T = int(os.environ.get("T", 4))
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
_, prec = make_cov(d, 1e4)
tgt = hx.targets.GaussianFull(precision=prec)
k_init, k_run = hx.random.split(hx.random.PRNGKey(0), 2)
x0 = hx.random.normal(k_init, (W, d), torch.float32)
keys = hx.random.split_device(k_run, 4 * T)
lib = _lib.load(); ctx = _rt.ctx(dev)
desc = tgt.descriptor(torch.float32, dev)
def run():
    X = x0.clone(); lp = tgt.log_prob(X)
    acc = torch.zeros(W, dtype=torch.int32, device=dev); flag = torch.zeros(1, dtype=torch.int32, device=dev)
    _lib.check(lib.hx_run_iterations(ctx, 0, 0, 0, 0, desc, _rt.ptr(X), _rt.ptr(lp), n, eps, L, _rt.ptr(keys), T, None, None,
                                     _rt.ptr(acc), _rt.ptr(flag), _rt.stream(dev)))
    torch.cuda.synchronize()
    return X, lp, acc
for form in ("stepwise", "propagator"):
    hx.config.update("leapfrog_form", form)
    outs = [run() for _ in range(3)]
    for i in (1, 2):
        same = [torch.equal(a, b) for a, b in zip(outs[0], outs[i])]
        nd = (outs[0][0] != outs[i][0]).any(dim=1).sum().item()
        print(form, "run", i, "vs 0: X/lp/acc equal:", same, " rows differing:", nd, " accepts", outs[0][2].sum().item(), outs[i][2].sum().item(), flush=True)
