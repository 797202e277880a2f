"""CPU oracle for the h-emcee ensemble-Hamiltonian hot path.

TEST INFRASTRUCTURE ONLY.  This package is a NumPy restatement of the
reference algorithm (clarkmiyamoto/h-emcee, `src/hemcee/...`) plus a bit-exact
restatement of the `jax.random` primitives that path calls.  It exists so the
CUDA path can be checked draw-for-draw without JAX (which is not installed in
the build image nor on the GPU box).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
`--impl reference` legs may import it.  The product package
(`h-emcee_b200/hemcee_b200`) never imports, links or executes anything here.

Parity pin status (see DESIGN.md "Oracle"):
  * `jaxrng`  — pinned by Random123 / JAX known-answer vectors (tests/golden/rng_kat.json)
  * sampler   — pinned by replaying the printed outputs of the reference's own
                tutorial notebooks (docs/tutorials/quickstart.ipynb cell 8/10,
                docs/tutorials/adaptation.ipynb cell 5); the reference ships no
                golden vectors and JAX cannot be imported here, so draw-level
                parity against a live reference run is NOT available.
"""
