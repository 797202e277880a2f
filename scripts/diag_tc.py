"""Diagnostic (GPU): tensor-core walk half-move forms vs the exact SIMT path at c5-like scale."""
import sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "h-emcee_b200")]
import hemcee_b200 as hx
from bench import make_cov

from hemcee_b200 import _lib as _l0, _rt as _r0
import torch as _t0
_l0.load().hx_ctx_set_tuning(_r0.ctx(_t0.device("cuda", 0)), 3, int(os.environ.get("EKIND", 2)))   # energy operand kind
_l0.load().hx_ctx_set_tuning(_r0.ctx(_t0.device("cuda", 0)), 14, int(os.environ.get("KF16", 0)))   # fp16 hi/lo kicks / apply
d = int(os.environ.get("D", 1000)); n = int(os.environ.get("N", 8192)); L = 10; eps = 0.1
_, prec = make_cov(d, 1e4)
tgt = hx.targets.GaussianFull(precision=prec)
k_init, k_run = hx.random.split(hx.random.PRNGKey(0), 2)
x0 = hx.random.normal(k_init, (2 * n, d), torch.float32)
keys = hx.random.split(k_run, 8)
for regime in ("N(0,I)", "burned"):
    G1, G2 = x0[:n].contiguous(), x0[n:].contiguous()
    if regime == "burned":
        hx.config.update("precision", "bf16x3"); hx.config.update("leapfrog_form", "stepwise")
        s = hx.HamiltonianEnsembleSampler(2 * n, d, tgt, step_size=eps, L=L, adapt_inital_step_size=False, adapt_step_size=False,
                                          adapt_length=False, verbose=False)
        s.run_mcmc(keys[5], x0, 30, warmup=0)
        X = s.last_state
        G1, G2 = X[:n].contiguous(), X[n:].contiguous()
    LP1 = tgt.log_prob(G1)
    res = {}
    for mode, form in (("exact", "auto"), ("bf16x3", "stepwise"), ("bf16x3", "propagator"), ("f16f8", "stepwise"), ("f16f8", "propagator"), ("tf32x3", "stepwise"), ("tf32x3", "propagator")):
        hx.config.update("precision", mode); hx.config.update("leapfrog_form", form)
        a = hx.hmc_walk_move(G1, G2, eps, keys[0], tgt, tgt, L, LP1)
        b = hx.hmc_walk_move(G1, G2, eps, keys[0], tgt, tgt, L, LP1)
        det = all(torch.equal(x, y) for x, y in zip(a, b))
        # with a prefetch hint in flight
        hx.moves.walk_prefetch(keys[0], n, 0, n)
        c0 = hx.hmc_walk_move(G1, G2, eps, keys[1], tgt, tgt, L, LP1)
        c = hx.hmc_walk_move(G1, G2, eps, keys[0], tgt, tgt, L, LP1)
        pre = all(torch.equal(x, y) for x, y in zip(a, c))
        res[(mode, form)] = a
        q, la, pp, lp = a
        qe, lae, ppe, lpe = res[("exact", "auto")]
        sc = qe.abs().max().item()
        print(f"[{regime}] {mode:7s} {form:10s} det={det} prefetch_same={pre}  |dq|/scale {(q-qe).abs().max().item()/sc:.2e}  |dS| rel {(pp-ppe).abs().max().item()/ppe.abs().max().item():.2e}"
              f"  dlp max {(lp-lpe).abs().max().item():.3e} rms {(lp-lpe).pow(2).mean().sqrt().item():.3e}"
              f"  dla max {(la-lae).abs().max().item():.3e} rms {(la-lae).pow(2).mean().sqrt().item():.3e}  mean acc {la.exp().mean().item():.4f}  lp mean {lp.mean().item():.2f}")
    # ---- f64 truth from the same f32 draws: S0 = P0 C in f64, exact f64 integrator, f64 energies
    from hemcee_b200 import _lib, _rt
    dev = G1.device
    P0 = hx.random.normal(keys[0], (n, n), torch.float32)
    X2d = G2.double()
    C = (X2d - X2d.mean(0)) / np.sqrt(n)
    S0 = P0.double() @ C
    del P0
    hx.config.update("enable_x64", True); hx.config.update("precision", "exact")
    tg64 = hx.targets.GaussianFull(precision=prec)
    q = G1.double().clone(); S = S0.clone(); dK = torch.empty(n, dtype=torch.float64, device=dev)
    _lib.check(_lib.load().hx_leapfrog_walk(_rt.ctx(dev), 1, tg64.descriptor(torch.float64, dev), _rt.ptr(q), _rt.ptr(S), _rt.ptr(C.contiguous()),
                                            n, n, eps, L, _rt.ptr(dK), _rt.stream(dev)), "leap")
    lp0 = tg64.log_prob(G1.double()); lpL = tg64.log_prob(q)
    dH = (lp0 - lpL) + dK
    la_t = torch.minimum(torch.zeros_like(dH), -dH)
    hx.config.update("enable_x64", False)
    print(f"[{regime}] truth mean acc {la_t.exp().mean().item():.4f}  lp mean {lpL.mean().item():.2f}   (lp0 f32 vs f64 rms {(LP1.double()-lp0).pow(2).mean().sqrt().item():.2e})")
    sc = q.abs().max().item()
    for kf, a in res.items():
        qq, la, pp, lp = [t.double() for t in a]
        print(f"   vs truth {kf[0]:7s} {kf[1]:10s} |dq|/scale {(qq-q).abs().max().item()/sc:.2e} |dS| rel {(pp-S).abs().max().item()/S.abs().max().item():.2e} "
              f"dlp rms {(lp-lpL).pow(2).mean().sqrt().item():.3e} dla max {(la-la_t).abs().max().item():.3e} rms {(la-la_t).pow(2).mean().sqrt().item():.3e} mean {(la-la_t).mean().item():+.3e}")
