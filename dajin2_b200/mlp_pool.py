"""Worker processes for the reference's lbfgs MLP (``mutation_processing/anomaly_detector.py:61-94``).

Loci selection trains one scikit-learn ``MLPClassifier`` per mutation type on 2 x (10 000 + L) synthetic points — a
fixed ~0.3-1.2 s of single-threaded host arithmetic per fit that does not depend on the number of reads.  The fit is
chaotic in its floating-point summation order (profiles/r01_mlp_stability.log: shuffling the training rows or a 1e-13
perturbation of the inputs changes the predicted candidate loci), so it cannot be re-implemented on the device and
stay bit-exact with the reference; it is therefore kept on scikit-learn, the reference's own dependency, with the
BLAS thread count pinned to 1 exactly as the reference pins it (``main.py:5-9``).  Two things are added here:
``fastmlp.FastMLPClassifier`` evaluates scikit-learn's loss/gradient with the same floating-point operations in fused
C passes (4-5x faster, checked against scikit-learn's own function in every fit), and the three fits of an allele
(and the fits of further alleles / samples) run side by side in a small pool of spawned workers while the GPU and
the calling thread carry on.

``DAJIN2_B200_MLP_WORKERS=0`` keeps the fits in the calling process.
"""

from __future__ import annotations

import atexit
import os
import pickle
import queue
import struct
import subprocess
import sys
import threading
from concurrent.futures import Future, ThreadPoolExecutor
from pathlib import Path

import numpy as np

#: what the most recent fits did (bench.py prints it): dicts {"n_iter", "fast_path", "seconds"}; bounded
STATS: list = []


def _note(stats: dict) -> None:
    STATS.append(stats)
    if len(STATS) > 256:
        del STATS[:128]


def fit_predict(sample: np.ndarray, control: np.ndarray, threshold: float) -> np.ndarray:
    """Candidate loci proposed by the reference's MLP; the training set is rebuilt exactly as the reference does."""
    return fit_predict_stats(sample, control, threshold)[0]


#: fits the caller expects to have in flight at once (3 per allele / sample in flight): pipeline.run_many raises it so
#: that the element-wise passes of each fit take fewer threads and the fits do not fight for the cores
expected_fits = 3


def threads_per_fit() -> int:
    """Threads of the element-wise passes of one fit: the cores this process may use, shared by the expected fits,
    never more than ``fastmlp.default_threads()``.  The thread count cannot change a bit of the result."""
    from .fastmlp import default_threads

    usable = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 4)
    return max(1, min(default_threads(), (usable - 2) // max(1, expected_fits)))


MLP_KW = dict(solver="lbfgs", alpha=1e-5, hidden_layer_sizes=(5, 2), random_state=1, max_iter=1000)


def training_set(control: np.ndarray, threshold: float):
    """The training set of ``detect_anomalies`` (``anomaly_detector.py:61-88``), rebuilt exactly as the reference does."""
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
    # the reference passes the labels as a Python list ([0] * n + [1] * n); the same labels as an array spare
    # scikit-learn's input validation the element-wise conversion (several ms per fit) and change nothing else
    y = np.repeat(np.array([0, 1]), n_rand + n_ctrl)
    return X, y


def fit_predict_stats(sample: np.ndarray, control: np.ndarray, threshold: float, threads: int | None = None,
                      defer_checks: bool = False):
    """``fit_predict`` plus what the fit did: (mask, {"n_iter", "fast_path", "seconds", "threads"}).  With
    ``defer_checks`` the fit does not compare itself with scikit-learn's evaluation; the points and fused results to
    compare are returned as ``stats["checks"]`` (see ``check_points``)."""
    import time

    if threads:
        from .fastmlp import load_host_lib

        load_host_lib().fm_set_threads(int(threads))
    t0 = time.perf_counter()
    import warnings

    from .fastmlp import FastMLPClassifier  # scikit-learn's MLPClassifier with a fused, bit-identical loss/gradient

    X, y = training_set(control, threshold)
    clf = FastMLPClassifier(**MLP_KW)
    clf.defer_checks = bool(defer_checks)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")  # ConvergenceWarning, as the reference silences it (utils/config.py:79-83)
        clf.fit(X, y)
    mask = clf.predict(np.array([control, sample]).T) == 1
    stats = {"n_iter": int(getattr(clf, "n_iter_", -1)), "fast_path": bool(getattr(clf, "fast_path_used_", False)),
             "seconds": time.perf_counter() - t0, "threads": int(threads or 0)}
    if defer_checks:
        stats["checks"] = clf.deferred_checks_ if clf.fast_path_used_ else []
    return mask, stats


def check_points(control: np.ndarray, threshold: float, points) -> list:
    """scikit-learn's own loss and gradient at ``points`` (packed parameter vectors; ``None`` = the initial ones) of
    the model ``fit_predict_stats`` fits for this control profile and threshold: [(packed, loss, grad), ...]."""
    import warnings

    from .fastmlp import ProbeMLPClassifier

    X, y = training_set(control, threshold)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        probe = ProbeMLPClassifier(points=list(points), **MLP_KW).fit(X, y)
    return probe.points_


def workers() -> int:
    raw = os.environ.get("DAJIN2_B200_MLP_WORKERS")
    if raw is not None:
        return max(0, int(raw))
    return max(3, min(16, (os.cpu_count() or 4) - 2))


# ---------------------------------------------------------------------------------------------------------
# the pool: N interpreter subprocesses running this file's worker loop, fed through pipes with pickled requests.
# (Not multiprocessing: its spawn / forkserver start methods re-import the parent's __main__ — bench.py, pytest —
# in every worker, and fork()ing a process that holds a CUDA context is not an option.)
# ---------------------------------------------------------------------------------------------------------


def _send(stream, obj) -> None:
    blob = pickle.dumps(obj, protocol=pickle.HIGHEST_PROTOCOL)
    stream.write(struct.pack("<Q", len(blob)))
    stream.write(blob)
    stream.flush()


def _recv(stream):
    head = stream.read(8)
    if len(head) < 8:
        raise EOFError("MLP worker closed its pipe")
    (n,) = struct.unpack("<Q", head)
    return pickle.loads(stream.read(n))


def _worker_main() -> None:
    stdin, stdout = sys.stdin.buffer, sys.stdout.buffer
    sys.stdout = sys.stderr  # stray prints must not corrupt the reply stream
    import sklearn.neural_network  # noqa: F401  (pay the import once, at start-up)

    _send(stdout, "ready")
    while True:
        try:
            req = _recv(stdin)
        except EOFError:
            return
        try:
            if req[0] == "check":
                _send(stdout, ("ok", check_points(*req[1:])))
            else:
                _send(stdout, ("ok", fit_predict_stats(*req[1:])))
        except BaseException as exc:  # noqa: BLE001 - reported to the parent
            _send(stdout, ("error", repr(exc)))


class _Worker:
    def __init__(self):
        env = dict(os.environ)
        for v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
            env[v] = "1"  # as the reference runs a sample (main.py:5-9): the fit is sensitive to the summation order
        env["CUDA_VISIBLE_DEVICES"] = ""
        env["OMP_WAIT_POLICY"] = "passive"  # idle threads of one fit must not spin on cores other fits need
        root = str(Path(__file__).resolve().parent.parent)
        env["PYTHONPATH"] = root + os.pathsep + env.get("PYTHONPATH", "")
        self.proc = subprocess.Popen([sys.executable, "-m", "dajin2_b200.mlp_pool", "--worker"], stdin=subprocess.PIPE,
                                     stdout=subprocess.PIPE, env=env, cwd=root)
        self.ready = False

    def call(self, args):
        if not self.ready:
            if _recv(self.proc.stdout) != "ready":
                raise RuntimeError("MLP worker failed to start")
            self.ready = True
        _send(self.proc.stdin, args)
        status, payload = _recv(self.proc.stdout)
        if status != "ok":
            raise RuntimeError(f"MLP worker: {payload}")
        return payload

    def close(self):
        try:
            self.proc.stdin.close()
            self.proc.wait(timeout=2)
        except Exception:  # noqa: BLE001
            self.proc.kill()


class _Pool:
    """n worker interpreters.  While fits with deferred self-checks are in flight (``submit(defer=True)``), up to three
    of the workers are kept free of fits so that scikit-learn's loss / gradient at the first and the last point of
    those fits is evaluated beside them: the fit's two self-checks leave its critical path.  Without such fits every
    worker takes fits."""

    def __init__(self, n: int):
        self.n = n
        self.all = [_Worker() for _ in range(n)]  # start every interpreter now: their imports run concurrently
        self.free = list(self.all)
        self.cv = threading.Condition()
        self.n_checkers = 3 if n >= 9 else (2 if n >= 6 else 0)
        self.deferred_in_flight = 0
        self.threads = ThreadPoolExecutor(max_workers=n, thread_name_prefix="dj-mlp")
        self.check_threads = ThreadPoolExecutor(max_workers=max(1, self.n_checkers), thread_name_prefix="dj-mlp-check")

    def _acquire(self, for_check: bool) -> _Worker:
        with self.cv:
            while True:
                reserve = 0 if for_check or self.deferred_in_flight == 0 else self.n_checkers
                if len(self.free) > reserve:
                    return self.free.pop()
                self.cv.wait()

    def _release(self, w: _Worker) -> None:
        with self.cv:
            self.free.append(w)
            self.cv.notify_all()

    def _check(self, control, threshold, points):
        w = self._acquire(True)
        try:
            return w.call(("check", control, threshold, points))
        finally:
            self._release(w)

    def _run(self, args, verdict: Future | None):
        first = None
        if verdict is not None:  # scikit-learn's evaluation of the initial point starts beside the fit
            first = self.check_threads.submit(self._check, args[1], args[2], [None])
        w = self._acquire(False)
        try:
            mask, stats = w.call(("fit",) + args + (verdict is not None,))
        except BaseException as exc:
            if verdict is not None:
                self._deferred_done()
                verdict.set_exception(exc)
            raise
        finally:
            self._release(w)
        checks = stats.pop("checks", None)
        _note(stats)
        if verdict is not None:
            verdict.add_done_callback(lambda _v: self._deferred_done())
            if not stats.get("fast_path"):
                verdict.set_result(True)  # the fit ran on scikit-learn's own evaluation: nothing to compare
            else:
                last = self.check_threads.submit(self._check, args[1], args[2], [checks[1][0]])

                def settle(_done, first=first, last=last, checks=checks):
                    try:
                        from .fastmlp import _same_bits

                        ref = [first.result()[0], last.result()[0]]
                        ok = all(np.array_equal(r[0].view(np.uint64), c[0].view(np.uint64))
                                 and _same_bits((r[1], r[2]), (c[1], c[2])) for r, c in zip(ref, checks))
                        verdict.set_result(bool(ok))
                    except BaseException as exc:  # noqa: BLE001
                        verdict.set_exception(exc)

                last.add_done_callback(settle)
        return mask

    def _deferred_done(self) -> None:
        with self.cv:
            self.deferred_in_flight -= 1
            self.cv.notify_all()

    def submit(self, args, defer: bool = False) -> Future:
        verdict: Future | None = Future() if (defer and self.n_checkers) else None
        if verdict is not None:
            with self.cv:
                self.deferred_in_flight += 1
        fut = self.threads.submit(self._run, args, verdict)
        fut.verdict = verdict
        return fut

    def close(self):
        self.threads.shutdown(wait=False, cancel_futures=True)
        self.check_threads.shutdown(wait=False, cancel_futures=True)
        for w in self.all:
            w.close()


_pool: _Pool | None = None


def _shutdown() -> None:
    global _pool
    if _pool is not None:
        _pool.close()
        _pool = None


#: cleared when a deferred comparison fails on this host: fits then check themselves inline again (and fall back to
#: scikit-learn's evaluation by themselves)
deferred_checks_ok = True


def submit(sample: np.ndarray, control: np.ndarray, threshold: float, defer_checks: bool = False) -> Future:
    """Start one fit; ``.result()`` gives the boolean candidate mask.  With ``defer_checks`` (and a pool with checker
    workers) the fit's comparisons with scikit-learn's own evaluation run in other workers and the caller MUST call
    ``verified(future)`` before it lets a result depend on the mask for good; False means: redo without deferral."""
    global _pool
    n = workers()
    if n == 0:
        fut: Future = Future()
        fut.verdict = None
        try:
            # as the pool's workers (and the reference, main.py:5-9): BLAS pinned to one thread for the fit, whose
            # result depends on the summation order
            try:
                from threadpoolctl import threadpool_limits
            except ImportError:  # the environment variables of the process decide
                import contextlib

                threadpool_limits = lambda **kw: contextlib.nullcontext()  # noqa: E731
            with threadpool_limits(limits=1, user_api="blas"):
                mask, stats = fit_predict_stats(sample, control, threshold, threads_per_fit())
            _note(stats)
            fut.set_result(mask)
        except BaseException as exc:  # noqa: BLE001 - handed to the caller through the future
            fut.set_exception(exc)
        return fut
    if _pool is None or _pool.n != n:
        _shutdown()
        _pool = _Pool(n)
        atexit.register(_shutdown)
    return _pool.submit((np.ascontiguousarray(sample, dtype=np.float64),
                         np.ascontiguousarray(control, dtype=np.float64), float(threshold), threads_per_fit()),
                        defer=defer_checks and deferred_checks_ok)


def verified(fut: Future) -> bool:
    """Outcome of the deferred comparisons of one fit (True when nothing was deferred).  Waits for them."""
    global deferred_checks_ok
    fut.result()
    verdict = getattr(fut, "verdict", None)
    if verdict is None:
        return True
    ok = bool(verdict.result())
    if not ok:
        deferred_checks_ok = False
    return ok


def warm_up() -> None:
    """Start the workers and import scikit-learn in each of them (first use otherwise pays ~1-2 s)."""
    n = workers()
    if n == 0:
        return
    probe = np.linspace(0.0, 3.0, 256)  # a real (small) fit: imports, the C helper's build and its thread team
    for f in [submit(probe, probe * 0.5, 0.1) for _ in range(n)]:
        f.result()


if __name__ == "__main__":
    if "--worker" in sys.argv:
        _worker_main()
