"""FastMLPClassifier must reproduce scikit-learn's MLPClassifier bit for bit (reference: ``detect_anomalies``,
src/DAJIN2/core/preprocess/mutation_processing/anomaly_detector.py:61-94)."""
import os
import warnings

import numpy as np
import pytest

for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
    os.environ.setdefault(_v, "1")

from sklearn.neural_network import MLPClassifier  # noqa: E402

from dajin2_b200 import fastmlp  # noqa: E402

KW = dict(solver="lbfgs", alpha=1e-5, hidden_layer_sizes=(5, 2), random_state=1, max_iter=1000)


def training_set(n_loci: int, seed: int, threshold: float = 0.1):
    """The reference's training set (anomaly_detector.py:65-87) on a synthetic control profile."""
    g = np.random.default_rng(seed)
    control = np.abs(g.normal(0.8, 0.2, n_loci))
    sample = np.clip(control + g.normal(0, 0.05, n_loci), 0, 100)
    sample[[n_loci // 5, n_loci // 3, n_loci // 2]] += [25, 15, 10]
    control = np.minimum(control, sample)
    rng = np.random.default_rng(seed=1)
    n_rand, n_ctrl = 10_000, len(control)
    base = rng.uniform(0, 100, n_rand)
    base_err = np.clip(base + rng.uniform(0, threshold, n_rand), 0, 100)
    base_mut = np.clip(base + rng.uniform(threshold, 100, n_rand), 0, 100)
    ctrl_err = np.clip(control + rng.uniform(0, threshold, n_ctrl), 0, 100)
    ctrl_mut = np.clip(control + rng.uniform(threshold, 100, n_ctrl), 0, 100)
    err = np.concatenate([np.array([base, base_err]).T, np.array([control, ctrl_err]).T], axis=0)
    mut = np.concatenate([np.array([base, base_mut]).T, np.array([control, ctrl_mut]).T], axis=0)
    X = np.concatenate([err, mut], axis=0)
    y = [0] * (n_rand + n_ctrl) + [1] * (n_rand + n_ctrl)
    return X, y, np.array([control, sample]).T


def _fit(cls, X, y):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return cls(**KW).fit(X, y)


@pytest.mark.parametrize("n_loci,seed", [(300, 9), (1000, 7), (2832, 8)])
def test_same_parameters_and_predictions(n_loci, seed):
    X, y, probe = training_set(n_loci, seed)
    ref, fast = _fit(MLPClassifier, X, y), _fit(fastmlp.FastMLPClassifier, X, y)
    for a, b in zip(ref.coefs_ + ref.intercepts_, fast.coefs_ + fast.intercepts_):
        assert np.array_equal(a.view(np.uint64), b.view(np.uint64))
    assert ref.n_iter_ == fast.n_iter_ and ref.loss_ == fast.loss_
    assert np.array_equal(ref.predict(probe), fast.predict(probe))


def test_every_evaluation_matches_sklearn():
    """Loss and gradient of the fused evaluation equal scikit-learn's at every iterate of a fit, bit for bit."""
    X, y, _ = training_set(500, 3)
    mismatches = []

    class Checked(fastmlp.FastMLPClassifier):
        def _loss_grad_lbfgs(self, packed, X, y, sw, act, deltas, cg, ig):
            if getattr(self, "_fm", None) is None:
                self._fast_setup(X, y)
            ref = MLPClassifier._loss_grad_lbfgs(self, packed.copy(), X, y, sw, act, deltas, cg, ig)
            got = self._fast_loss_grad(packed, X)
            if not fastmlp._same_bits(ref, got):
                mismatches.append(len(mismatches))
            return ref

        def _fit_lbfgs(self, *a):
            return MLPClassifier._fit_lbfgs(self, *a)

    m = _fit(Checked, X, y)
    assert m.n_iter_ > 10 and not mismatches


def test_fallback_when_the_fused_path_does_not_apply():
    """Other layer sizes may take other BLAS paths: the self-check must notice and leave the fit to scikit-learn."""
    X, y, _ = training_set(400, 11)
    kw = dict(solver="lbfgs", alpha=1e-4, hidden_layer_sizes=(4, 3), random_state=3, max_iter=200)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref, got = MLPClassifier(**kw).fit(X, y), fastmlp.FastMLPClassifier(**kw).fit(X, y)
    for a, b in zip(ref.coefs_ + ref.intercepts_, got.coefs_ + got.intercepts_):
        assert np.array_equal(a, b)


def test_fast_path_is_taken_here():
    X, y, _ = training_set(300, 5)
    assert _fit(fastmlp.FastMLPClassifier, X, y).fast_path_used_


def test_fit_predict_equals_the_reference_call():
    """mlp_pool.fit_predict (labels as an array, fused evaluation) proposes exactly the candidate loci of the reference's
    own call sequence (anomaly_detector.py:61-94, restated in the oracle with scikit-learn's MLPClassifier)."""
    import numpy as np

    from dajin2_b200 import mlp_pool
    from oracle import dajin2_oracle as orc

    rng = np.random.default_rng(12)
    for L, thr in ((700, 0.1), (1500, 10.0)):
        control = np.abs(rng.normal(0.8, 0.15, L))
        sample = control + rng.normal(0, 0.05, L)
        sample[[L // 5, L // 3, L // 2]] += [25, 15, 60]
        sample = np.clip(sample, 0, 100)
        control = np.minimum(control, sample)
        got = mlp_pool.fit_predict(sample, control, thr)
        assert set(np.flatnonzero(got).tolist()) == set(orc.mlp_candidates(sample, control, thr))


def test_fast_validation_leaves_the_same_estimator():
    """FastMLPClassifier fitted on array labels (short-cut input validation) == MLPClassifier fitted on the reference's
    list labels: parameters, iteration count, loss, classes, predictions and probabilities."""
    import sys
    import warnings
    from pathlib import Path

    import numpy as np
    from sklearn.neural_network import MLPClassifier

    sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "scripts"))
    from mlp_case import case

    from dajin2_b200.fastmlp import FastMLPClassifier

    X, y = case(900, 4)
    kw = dict(solver="lbfgs", alpha=1e-5, hidden_layer_sizes=(5, 2), random_state=1, max_iter=1000)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = MLPClassifier(**kw).fit(X, y)
        got = FastMLPClassifier(**kw).fit(X, np.array(y))
        row_major = FastMLPClassifier(**kw).fit(np.ascontiguousarray(X), np.array(y))
    assert X.flags.f_contiguous and not X.flags.c_contiguous  # what the reference's construction yields
    for m in (got, row_major):
        assert m.fast_path_used_ and m.fast_validation_used_
        assert all(np.array_equal(a, b) for a, b in zip(ref.coefs_ + ref.intercepts_, m.coefs_ + m.intercepts_))
    assert (ref.n_iter_, ref.loss_, ref.n_features_in_, ref.n_outputs_) == (got.n_iter_, got.loss_, got.n_features_in_, got.n_outputs_)
    assert ref.classes_.tolist() == got.classes_.tolist() == [0, 1]
    probe = np.random.default_rng(0).uniform(0, 100, (3000, 2))
    assert np.array_equal(ref.predict(probe), got.predict(probe))
    assert np.array_equal(ref.predict_proba(probe), got.predict_proba(probe))
    # labels outside {0, 1} or non-finite inputs take scikit-learn's own validation
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        other = FastMLPClassifier(**kw).fit(X, np.array(y) + 3)
    assert other.classes_.tolist() == [3, 4] and not other.fast_validation_used_


def test_vector_exp_log_return_libms_bits():
    """hostc/vmath.c: the four-wide exp / log built on the tables of the loaded libm agree with libm bit for bit on
    4 x 10^6 arguments each (ranges of the MLP, special-case boundaries, random bit patterns); with tables that are
    not libm's the self test fails, the pass stays off and fits go through libm's own calls -- same result."""
    import warnings

    import numpy as np
    import pytest
    from sklearn.neural_network import MLPClassifier

    from dajin2_b200 import fastmlp

    lib = fastmlp.load_host_lib()
    tables = fastmlp._libm_tables()
    if tables is None or lib.fm_vmath_init(*(t.ctypes.data for t in tables), 10) < 0:
        pytest.skip("no AVX2 + FMA, or a libm whose exp / log tables are laid out differently")
    assert lib.fm_vmath_init(*(t.ctypes.data for t in tables), 1_000_000) == 0 and lib.fm_vmath_ready() == 1
    if lib.fm_get_avx512():  # the eight-wide form (vmath512.c) on the same terms
        assert lib.fm512_vmath_init(*(t.ctypes.data for t in tables), 500_000) == 0 and lib.fm512_vmath_ready() == 1
    rng = np.random.default_rng(2)
    X = np.concatenate([rng.uniform(0, 100, (600, 2)), rng.uniform(0, 3, (600, 2))])
    y = np.repeat(np.array([0, 1]), 600)
    kw = dict(solver="lbfgs", alpha=1e-5, hidden_layer_sizes=(5, 2), random_state=1, max_iter=300)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = MLPClassifier(**kw).fit(X, y.tolist())
        on = fastmlp.FastMLPClassifier(**kw).fit(X, y)
        wrong = [t.copy() for t in tables]
        wrong[1][5] ^= np.uint64(1)  # one bit of one table entry
        assert lib.fm_vmath_init(*(t.ctypes.data for t in wrong), 100_000) > 0 and lib.fm_vmath_ready() == 0
        if lib.fm_get_avx512():
            four = fastmlp.FastMLPClassifier(**kw).fit(X, y)  # eight-wide pass still on, four-wide one off
            assert lib.fm512_vmath_init(*(t.ctypes.data for t in wrong), 100_000) > 0 and lib.fm512_vmath_ready() == 0
        off = fastmlp.FastMLPClassifier(**kw).fit(X, y)
        assert lib.fm_vmath_init(*(t.ctypes.data for t in tables), 100_000) == 0
        if lib.fm_get_avx512():
            assert lib.fm512_vmath_init(*(t.ctypes.data for t in tables), 100_000) == 0
            lib.fm_set_avx512(0)
            narrow = fastmlp.FastMLPClassifier(**kw).fit(X, y)  # AVX-512 off altogether: the four-wide pass
            lib.fm_set_avx512(1)
            for m in (four, narrow):
                assert m.fast_path_used_
                assert all(np.array_equal(a, b) for a, b in zip(ref.coefs_ + ref.intercepts_, m.coefs_ + m.intercepts_))
    for m in (on, off):
        assert m.fast_path_used_
        assert all(np.array_equal(a, b) for a, b in zip(ref.coefs_ + ref.intercepts_, m.coefs_ + m.intercepts_))


def test_avx512_passes_return_the_avx2_bits():
    """hostc/fastmlp512.c: forward activations and the packed gradient of the 2-5-2 model, AVX-512 form against the
    AVX2 form, bit for bit -- row counts around the eight-row trips, column-major X, dead units (zeros and negative
    zeros behind the relu masks)."""
    import numpy as np
    import pytest

    from dajin2_b200 import fastmlp

    lib = fastmlp.load_host_lib()
    P = fastmlp._ptr
    lib.fm_set_avx512(1)
    if not lib.fm_get_avx512():
        pytest.skip("no AVX-512 on this machine")

    def run(on, X, packed, d3):
        n = X.shape[0]
        lib.fm_set_avx512(on)
        a1, a2 = np.full((n, 5), np.nan), np.full((n, 2), np.nan)
        sr, sc = X.strides[0] // 8, X.strides[1] // 8
        assert lib.fm_forward_packed(P(X), sr, sc, n, 2, 5, 2, P(packed), P(a1), P(a2)) == 0
        d2, d1, g3, grad = np.empty((n, 2)), np.empty((n, 5)), a2.T @ d3, np.empty(packed.shape[0])
        assert lib.fm_backward_packed(P(X), sr, sc, n, 2, 5, 2, P(packed), P(a1), P(a2), P(d3), P(d2), P(d1), P(g3), 0.5,
                                      1e-5, P(grad)) == 0
        return a1, a2, grad

    rng = np.random.default_rng(0)
    try:
        for n in (1, 7, 8, 9, 16, 17, 33, 1001, 30003):
            for trial in range(3):
                X = rng.uniform(0, 100, (n, 2))
                if trial == 2:
                    X = np.asfortranarray(X)
                packed = rng.normal(0, 0.7, 2 * 5 + 5 * 2 + 2 + 5 + 2 + 1)
                if trial == 1:
                    packed[10:12] = 0.0
                    packed[:5] = -np.abs(packed[:5])
                d3 = rng.normal(0, 1, (n, 1))
                for a, b in zip(run(0, X, packed, d3), run(1, X, packed, d3)):
                    assert np.array_equal(a.view(np.uint64), b.view(np.uint64)), (n, trial)
    finally:
        lib.fm_set_avx512(1)


def test_deferred_checks_leave_the_fit_and_the_verdict_unchanged():
    """defer_checks: the fit does not call scikit-learn's evaluation itself; ProbeMLPClassifier (scikit-learn's own
    _loss_grad_lbfgs after scikit-learn's own initialisation) returns the same bits at the first and the last
    evaluated point as the fused evaluation recorded, and the fitted parameters are those of the self-checking fit."""
    X, y, _ = training_set(700, 3)
    y = np.asarray(y)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        a = fastmlp.FastMLPClassifier(**KW)
        a.defer_checks = True
        a.fit(X, y)
        b = fastmlp.FastMLPClassifier(**KW).fit(X, y)
        assert a.fast_path_used_ and len(a.deferred_checks_) == 2
        assert all(np.array_equal(p.view(np.uint64), q.view(np.uint64)) for p, q in zip(a.coefs_ + a.intercepts_,
                                                                                        b.coefs_ + b.intercepts_))
        probe = fastmlp.ProbeMLPClassifier(points=[None, a.deferred_checks_[1][0]], **KW).fit(X, y)
    for (p0, l0, g0), (p1, l1, g1) in zip(a.deferred_checks_, probe.points_):
        assert np.array_equal(p0.view(np.uint64), p1.view(np.uint64))
        assert fastmlp._same_bits((l0, g0), (l1, g1))
    # a wrong fused value is caught by the comparison
    assert not fastmlp._same_bits((a.deferred_checks_[1][1] + 1e-16 * abs(a.deferred_checks_[1][1]) + 5e-324,
                                   a.deferred_checks_[1][2]), (probe.points_[1][1], probe.points_[1][2]))


def test_pool_defers_checks_to_other_workers(monkeypatch):
    """mlp_pool.submit(defer_checks=True): same masks as the self-checking fits, verdict True; a tampered fused result
    gives verdict False and switches deferral off for the process."""
    from dajin2_b200 import mlp_pool

    monkeypatch.setenv("DAJIN2_B200_MLP_WORKERS", "6")
    g = np.random.default_rng(5)
    control = np.abs(g.normal(0.5, 0.3, 600))
    sample = control.copy()
    sample[[100, 200]] = [30.0, 12.0]
    plain = [mlp_pool.submit(sample, control, th) for th in (0.1, 0.2)]
    deferred = [mlp_pool.submit(sample, control, th, defer_checks=True) for th in (0.1, 0.2)]
    for p, d in zip(plain, deferred):
        assert np.array_equal(p.result(), d.result())
        assert mlp_pool.verified(p) and mlp_pool.verified(d)
        assert d.verdict is not None and p.verdict is None
    assert mlp_pool.deferred_checks_ok
    real = fastmlp._same_bits
    try:
        fastmlp._same_bits = lambda ref, got: False
        bad = mlp_pool.submit(sample, control, 0.1, defer_checks=True)
        assert not mlp_pool.verified(bad)
        assert not mlp_pool.deferred_checks_ok
        again = mlp_pool.submit(sample, control, 0.1, defer_checks=True)  # not deferred any more
        assert again.verdict is None and np.array_equal(again.result(), plain[0].result())
    finally:
        fastmlp._same_bits = real
        mlp_pool.deferred_checks_ok = True
        mlp_pool._shutdown()


def test_one_call_evaluation_returns_the_three_call_bits():
    """fm_eval (one C call per evaluation; the two matrix-vector products and the dots made by numpy's own OpenBLAS
    through its cblas entry points) is in use here, and a fit through it equals a fit through the three-call path
    (products through numpy) in every parameter bit and in the number of evaluations."""
    X, y, _ = training_set(900, 4)
    y = np.asarray(y)
    assert fastmlp._numpy_blas() is not None
    saved = fastmlp._BLAS
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            one = fastmlp.FastMLPClassifier(**KW).fit(X, y)
            fastmlp._BLAS = None
            three = fastmlp.FastMLPClassifier(**KW).fit(X, y)
    finally:
        fastmlp._BLAS = saved
    assert one.fast_path_used_ and three.fast_path_used_ and one.n_iter_ == three.n_iter_
    assert all(np.array_equal(p.view(np.uint64), q.view(np.uint64))
               for p, q in zip(one.coefs_ + one.intercepts_, three.coefs_ + three.intercepts_))
    assert float(one.loss_) == float(three.loss_)
