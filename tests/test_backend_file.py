"""FileBackend: the in-memory chain store written to / restored from one file with the dataset and attribute names of the
reference's HDF group (reference src/hemcee/backend/hdf.py:99-128); slicing semantics of backend/backend.py:52-71."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "h-emcee_b200"))


def test_file_backend_round_trip(tmp_path):
    from hemcee_b200.backend import Backend, FileBackend
    rng = np.random.default_rng(0)
    W, d = 6, 3
    fn = str(tmp_path / "chain.npz")
    b = FileBackend(fn)
    b.reset(W, d)
    warm = rng.standard_normal((4, W, d)).astype(np.float32)
    b.save_slice(warm, warm.sum(-1), np.arange(W), 0)
    b.mark_warmup_end()
    main = rng.standard_normal((5, W, d)).astype(np.float32)
    b.save_slice(main, main.sum(-1), np.ones(W), 4)
    b.save()
    r = FileBackend.load(fn)
    assert (r.nwalkers, r.ndim, r.iteration, r.iteration_warmup, r.iteration_main) == (W, d, 9, 4, 5)
    np.testing.assert_array_equal(r.get_chain(), np.concatenate([warm, main]))
    np.testing.assert_array_equal(r.get_chain(discard=4, thin=2), main[::2])
    np.testing.assert_array_equal(r.get_chain(discard=4, flat=True), main.reshape(-1, d))
    np.testing.assert_array_equal(r.get_log_prob(discard=4), main.sum(-1))
    np.testing.assert_array_equal(r.accepted, np.arange(W) + 1.0)
    last, last_lp = r.get_last_sample()
    np.testing.assert_array_equal(last, main[-1])
    # same dataset / attribute names as the reference's HDF group
    with np.load(fn) as z:
        names = set(z.files)
    for k in ("chain", "log_prob", "accepted", "attrs/version", "attrs/nwalkers", "attrs/ndim", "attrs/iteration",
              "attrs/iteration_warmup", "attrs/iteration_main", "attrs/warmup_end_index"):
        assert f"mcmc/{k}" in names
    with pytest.raises(RuntimeError):
        r.save()                                   # loaded read-only, like HDFBackend.open


def test_file_backend_before_warmup_end(tmp_path):
    from hemcee_b200.backend import FileBackend
    fn = str(tmp_path / "c.npz")
    b = FileBackend(fn, name="run1")
    b.reset(4, 2)
    x = np.zeros((3, 4, 2), np.float32)
    b.save_slice(x, x.sum(-1), np.zeros(4), 0)
    b.save()
    r = FileBackend.load(fn, name="run1")
    assert r._warmup_end_index is None and r.iteration_warmup == 3 and r.iteration_main == 0
    with pytest.raises(KeyError):
        FileBackend.load(fn, name="mcmc")


def test_thinned_layout_maps_discard_and_thin_onto_stored_rows():
    """store='thinned' keeps iterations warmup, warmup + thin_by, ...; the backend still counts the iterations RUN and serves
    the reference's `v[discard::thin]` requests that land on stored rows (reference backend/backend.py:63-66)."""
    from hemcee_b200.backend import Backend
    b = Backend()
    b.reset(4, 2)
    warmup, thin_by, num = 10, 3, 5
    full = np.arange((warmup + num * thin_by) * 4 * 2, dtype=np.float32).reshape(-1, 4, 2)
    stored = full[warmup::thin_by]
    assert stored.shape[0] == num
    b.save_slice(stored, stored[:, :, 0], np.zeros(4), 0)
    b.iteration, b.iteration_main = warmup + num * thin_by, num * thin_by
    b.set_stored_layout(warmup, thin_by, last_sample=(full[-1], full[-1, :, 0]))
    np.testing.assert_array_equal(b.get_chain(discard=warmup, thin=thin_by), full[warmup::thin_by])
    np.testing.assert_array_equal(b.get_chain(discard=warmup + thin_by, thin=2 * thin_by), full[warmup + thin_by::2 * thin_by])
    with pytest.raises(ValueError):
        b.get_chain(discard=warmup + 1, thin=thin_by)     # iteration 11 is not a stored row
    with pytest.raises(ValueError):
        b.get_chain(discard=0, thin=thin_by)              # warm-up iterations were not stored
    with pytest.raises(ValueError):
        b.get_chain(discard=warmup, thin=2)               # not a multiple of the stored stride
    x, lp = b.get_last_sample()                           # the final iteration itself is not a stored row
    np.testing.assert_array_equal(x, full[-1])
    np.testing.assert_array_equal(b.get_autocorr_time.__self__.get_chain(discard=warmup, thin=thin_by, flat=True),
                                  full[warmup::thin_by].reshape(-1, 2))


def test_empty_store_returns_empty_arrays():
    from hemcee_b200.backend import Backend
    b = Backend()
    b.reset(4, 2)
    b.iteration = 7                                       # iterations were run but none was stored (num_samples = 0)
    assert b.get_chain().shape == (0, 4, 2) and b.get_log_prob().shape == (0, 4)


def test_file_backend_autosave_keeps_completed_slices(tmp_path):
    """autosave=True: the file on disk holds every completed slice (what the reference's HDFBackend guarantees by appending in
    save_slice, hdf.py:155-170), so a run that dies later loses only the unfinished batch."""
    from hemcee_b200.backend import FileBackend
    fn = str(tmp_path / "auto.npz")
    b = FileBackend(fn, autosave=True)
    b.reset(4, 2)
    rng = np.random.default_rng(0)
    c1, c2 = rng.standard_normal((3, 4, 2)).astype(np.float32), rng.standard_normal((2, 4, 2)).astype(np.float32)
    b.save_slice(c1, c1[:, :, 0], np.ones(4), 0)
    r = FileBackend.load(fn)
    assert r.iteration == 3 and r._warmup_end_index is None
    np.testing.assert_array_equal(r.get_chain(), c1)
    b.mark_warmup_end()
    b.save_slice(c2, c2[:, :, 0], np.ones(4), 3)
    r = FileBackend.load(fn)
    assert r.iteration == 5 and r.iteration_warmup == 3 and r.iteration_main == 2
    np.testing.assert_array_equal(r.get_chain(), np.concatenate([c1, c2]))
    np.testing.assert_array_equal(r.accepted, 2 * np.ones(4))
