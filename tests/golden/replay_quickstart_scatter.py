"""How much do the tuned integration length and the autocorrelation time of docs/tutorials/quickstart.ipynb scatter from one
random stream to the next?  The notebook (PRNGKey(2), unknown JAX build) printed L = 23, tau ~ 0.47; the oracle replays of the
same key give L = 20 (partitionable threefry) and L = 24 (legacy).  This script repeats the run with other seeds and records
(found step, ChEES T_bar, found L, tau).  `finalize` maps T_bar to L deterministically — L = ceil(T_bar (0.4 + 0.6 h) / step)
with h = halton(200001) = 0.5105 after 2 x 1e5 adapter updates (reference adaptation/chees.py:168-190) — so the scatter of L
is the scatter of T_bar, an exponential moving average (alpha = 0.9) of the ADAM-driven log T.
Run:  python tests/golden/replay_quickstart_scatter.py [seeds...]      (4 min per seed, one process each)
Writes tests/golden/quickstart_scatter.json."""
import json, os, sys, time
from multiprocessing import Pool
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))


def run(seed):
    from oracle import jaxrng as R, targets, sampler, autocorr
    dim, W = 3, 20
    keys = R.split(R.PRNGKey(seed), 2)
    init = R.normal(keys[0], (W, dim), np.float64)
    s = sampler.HamiltonianEnsembleSampler(W, dim, targets.make_iso_gaussian(), dtype=np.float64, verbose=False)
    samples = s.run_mcmc(keys[1], init, num_samples=10**4, warmup=10**5, thin_by=1)
    tau = autocorr.integrated_time(samples, quiet=True)
    ch = s.adapter_state.chees_state
    return dict(seed=seed, initial_step=float(s.step_size), found_step=float(s.found[0]), found_L=int(s.found[1]),
                T_bar=float(np.exp(ch.log_T_bar)), halton_final=float(ch.halton), chees_iteration=float(ch.iteration),
                tau=[float(t) for t in tau], acceptance=float(np.mean(s.diagnostics_main["acceptance_rate"])))


if __name__ == "__main__":
    seeds = [int(a) for a in sys.argv[1:]] or [0, 1, 2, 3, 4, 5]
    with Pool(min(len(seeds), os.cpu_count() or 1)) as p:
        out = p.map(run, seeds)
    print(json.dumps(out, indent=1))
    with open(os.path.join(os.path.dirname(__file__), "quickstart_scatter.json"), "w") as f:
        json.dump(out, f, indent=1)
