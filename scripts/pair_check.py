"""GPU check: CTA-pair (cta_group::2) projection kernel against the single-CTA kernel (move level and batch loop)."""
import sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "h-emcee_b200")]
import hemcee_b200 as hx
from hemcee_b200 import _lib, _rt
from bench import make_cov
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
lib = _lib.load(); ctx = _rt.ctx(dev); KNOB = int(os.environ.get("KNOB", 8))
for prec in ("bf16x3", "f16f8"):
    hx.config.update("precision", prec)
    for n, d, L in ((256, 300, 3), (1024, 1000, 5), (4096, 520, 4), (8192, 1000, 10)):
        _, P = make_cov(d, 100.0)
        tgt = hx.targets.GaussianFull(precision=P)
        k_init, k_run = hx.random.split(hx.random.PRNGKey(n + d), 2)
        x0 = hx.random.normal(k_init, (2 * n, d), torch.float32)
        G1, G2 = x0[:n].contiguous(), x0[n:].contiguous()
        LP1 = tgt.log_prob(G1)
        out = {}
        for pair in (0, 1):
            lib.hx_ctx_set_tuning(ctx, KNOB, pair)
            out[pair] = hx.hmc_walk_move(G1, G2, 0.02, k_run, tgt, tgt, L, LP1)
            torch.cuda.synchronize()
        dq = (out[0][0] - out[1][0]).abs().max().item() / out[0][0].abs().max().item()
        ds = (out[0][2] - out[1][2]).abs().max().item() / out[0][2].abs().max().item()
        print(f"{prec} move n={n} d={d}: |dq|/scale {dq:.2e} |dS| rel {ds:.2e} bit-equal {all(torch.equal(a, b) for a, b in zip(out[0], out[1]))}", flush=True)
        ch = {}
        for pair in (0, 1):
            lib.hx_ctx_set_tuning(ctx, KNOB, pair)
            s = hx.HamiltonianEnsembleSampler(2 * n, d, tgt, step_size=0.02, L=L, adapt_inital_step_size=False, adapt_step_size=False,
                                              adapt_length=False, verbose=False)
            ch[pair] = s.run_mcmc(k_run, x0, 3, warmup=0)
        ch0, ch1 = np.asarray(ch[0]), np.asarray(ch[1])
        print(f"{prec} chain n={n} d={d}: max diff {np.abs(ch0 - ch1).max() / np.abs(ch0).max():.2e} bit-equal {np.array_equal(ch0, ch1)}", flush=True)
lib.hx_ctx_set_tuning(ctx, KNOB, 0)
print("pair check done")
