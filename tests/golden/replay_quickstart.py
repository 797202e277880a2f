"""Replay of the reference's docs/tutorials/quickstart.ipynb cells 4-10 with the oracle.

The notebook's printed outputs are the only end-to-end numbers the reference
publishes for this path (JAX is not installable here, so a live run is not
possible).  Expected, from the notebook's stored output cells:
    Using inital step size of 1.6
    Found step size 0.10160634057813525 and integration length: 23
    acceptance_rate = [0.9992 0.9991 0.9993 0.9991 0.9994 0.9989 0.9988 0.9992 0.9991 0.9985
                       0.9993 0.9988 0.9991 0.9989 0.9991 0.9994 0.999  0.9988 0.9995 0.9985]
    tau = [0.4727686  0.4846164  0.46748046]
Run:  python tests/golden/replay_quickstart.py [warmup] [num_samples] [impl] [conditional]   (≈ minutes on one core)
Writes tests/golden/quickstart_replay_<impl>.json; with `conditional` the adapters are off and the notebook's found values
(step 0.10160634057813525, L = 23) are used from the start (main-phase dynamics given the tuned parameters; writes
quickstart_replay_conditional.json).
"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from oracle import jaxrng as R, moves, targets, sampler, autocorr

warmup = int(sys.argv[1]) if len(sys.argv) > 1 else 10**5
num = int(sys.argv[2]) if len(sys.argv) > 2 else 10**4
impl = sys.argv[3] if len(sys.argv) > 3 else "partitionable"
dim, W = 3, 20
key = R.PRNGKey(2)
keys = R.split(key, 2, impl)
init = R.normal(keys[0], (W, dim), np.float64, impl)
cond = len(sys.argv) > 4 and sys.argv[4] == "conditional"
if cond:
    s = sampler.HamiltonianEnsembleSampler(W, dim, targets.make_iso_gaussian(), step_size=0.10160634057813525, L=23,
                                           adapt_inital_step_size=False, adapt_step_size=False, adapt_length=False,
                                           dtype=np.float64, impl=impl, verbose=True)
else:
    s = sampler.HamiltonianEnsembleSampler(W, dim, targets.make_iso_gaussian(), dtype=np.float64, impl=impl, verbose=True)
t0 = time.time()
samples = s.run_mcmc(keys[1], init, num_samples=num, warmup=warmup, thin_by=1)
t1 = time.time()
acc = s.diagnostics_main["acceptance_rate"]
tau = autocorr.integrated_time(samples, quiet=True)
out = dict(impl=impl, warmup=warmup, num_samples=num, initial_step=float(s.step_size),
           found_step=float(s.found[0]), found_L=int(s.found[1]), acceptance_rate=[float(a) for a in acc],
           tau=[float(t) for t in tau], seconds=t1 - t0)
print(json.dumps(out, indent=1))
if cond and num == 10**4:
    with open(os.path.join(os.path.dirname(__file__), "quickstart_replay_conditional.json"), "w") as f:
        json.dump(out, f, indent=1)
elif warmup == 10**5 and num == 10**4:
    with open(os.path.join(os.path.dirname(__file__), f"quickstart_replay_{impl}.json"), "w") as f:
        json.dump(out, f, indent=1)
