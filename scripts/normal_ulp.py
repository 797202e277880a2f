"""GPU diagnostic: ulp distance of the device f32 normals from the oracle (numpy float32 restatement of XLA's erf_inv)."""
import sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "h-emcee_b200")]
import hemcee_b200 as hx
from oracle import jaxrng as R
key = R.PRNGKey(5)
got = hx.random.normal(key, (2000, 2001), torch.float32).cpu().numpy()
ref = R.normal(key, (2000, 2001), np.float32)
ulp = np.spacing(np.abs(ref).astype(np.float32))
e = np.abs(got.astype(np.float64) - ref) / ulp
print("n", got.size, "equal", np.mean(got == ref), "<=1ulp", np.mean(e <= 1), "<=2ulp", np.mean(e <= 2), "max ulp", e.max(), "at |z|", np.abs(ref.ravel()[e.argmax()]))
tail = np.abs(ref) > 2.7
print("tail (|z|>2.7): n", tail.sum(), "equal", np.mean(got[tail] == ref[tail]), "max ulp", e[tail].max())
small = np.abs(ref) < 1e-3
print("small (|z|<1e-3): n", small.sum(), "max ulp", e[small].max() if small.any() else None)
