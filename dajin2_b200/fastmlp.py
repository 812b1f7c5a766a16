"""A faster evaluation of scikit-learn's lbfgs MLP loss/gradient that is bit-identical to scikit-learn's own.

``detect_anomalies`` of the reference (``mutation_processing/anomaly_detector.py:61-94``) fits
``MLPClassifier(solver="lbfgs", alpha=1e-5, hidden_layer_sizes=(5, 2), random_state=1, max_iter=1000)`` on
2 x (10 000 + L) two-dimensional points, three times per allele: ~0.6 s of single-threaded host time that does not
depend on the number of reads and dominates a step once the read-scale passes run on the GPU.  The fit is chaotic in
its rounding (profiles/r01_mlp_stability.log), so anything short of the same floating-point operations in the same
order selects different candidate loci.  ``FastMLPClassifier`` therefore changes nothing about the model, the
initialisation or the optimiser (all inherited from scikit-learn / scipy); it only replaces the ~60 numpy calls of
one loss/gradient evaluation (``_loss_grad_lbfgs``) by the same arithmetic fused into three C passes
(``hostc/fastmlp.c``: hidden layers forward; logistic + log-loss + output delta; backward), about ten times faster:

  * the matrix products numpy hands to dgemm are single fused-multiply-add chains over the inner dimension
    (``scripts/micro/blas_order_probe.py``) and are restated as such; the two matrix x vector products of the output
    unit (dgemv: another order) are still made by numpy, exactly as scikit-learn's ``safe_sparse_dot`` makes them;
  * sums are numpy's pairwise summation or row-by-row accumulation, whichever numpy uses for that shape;
  * ``expit`` / ``xlogy`` are libm's ``exp`` / ``log`` as in scipy.special; these two take most of the time left and
    run on ``DAJIN2_B200_MLP_THREADS`` threads (element-wise, so the thread count cannot change a bit), each thread
    also summing the leaves of numpy's pairwise sums that lie in its rows;
  * the hidden layers' forward pass and the whole backward pass (deltas and all 27 reduction chains in one sweep) run on
    one thread with AVX2 + FMA, vector lanes being different output elements, never a split sum: spread over threads
    these passes were bounded by their arrays travelling between cores (``scripts/mlp_breakdown.py``,
    ``scripts/mlp_concurrency.py``: the longest of three concurrent fits 246 -> 195 ms on the 16-core B200 host).

Trust, but verify: every fit evaluates scikit-learn's original ``_loss_grad_lbfgs`` at the first and at the last
iterate and compares loss and gradient bit for bit; on any difference (another numpy / scipy / libm than the ones
this was written against) the fit is redone with scikit-learn's own evaluation, so the result is always
scikit-learn's.  ``tests/test_fastmlp.py`` compares complete fits.
"""

from __future__ import annotations

import copy
import ctypes
import os
import platform
import subprocess
from pathlib import Path

import numpy as np
from sklearn.neural_network import MLPClassifier

HERE = Path(__file__).resolve().parent
SOURCES = [HERE / "hostc" / name for name in ("fastmlp.c", "choice.c", "ingest.c", "midsvfix.c", "samscan.c")]  # one small host library
VMATH = HERE / "hostc" / "vmath.c"
AVX512 = HERE / "hostc" / "fastmlp512.c"  # compiled for AVX-512 on any x86-64 build machine, entered after a run-time check
VMATH512 = HERE / "hostc" / "vmath512.c"
HEADERS = [HERE / "hostc" / "vmath_lanes.h"]
LIB = HERE / "hostc" / "libdjhost.so"

_lib = None
_c_double_p = ctypes.POINTER(ctypes.c_double)


def build(force: bool = False) -> Path:
    """gcc -O3 without fast-math and without FMA contraction: the C code must round exactly like numpy does.  Only
    ``vmath.c`` (exp / log that return libm's bits) is compiled with ``-ffp-contract=fast``, as glibc itself is."""
    sources = SOURCES + [VMATH, AVX512, VMATH512]
    if force or not LIB.exists() or LIB.stat().st_mtime < max(src.stat().st_mtime for src in sources + HEADERS):
        tmp = LIB.with_suffix(".so.tmp%d" % os.getpid())
        flags = ["-O3", "-fPIC", "-fno-fast-math", "-fopenmp"]
        try:
            cpu = Path("/proc/cpuinfo").read_text()
            if " fma " in cpu:
                flags.append("-mfma")  # fma() inlines to vfmadd; without it libm's (exact, slow) fma is called
            if " avx2 " in cpu:
                flags.append("-mavx2")  # the loops over rows vectorise (rows are independent: same bits)
        except OSError:
            pass
        objs = []
        for src in sources:
            obj = LIB.parent / (src.name + ".%d.o" % os.getpid())
            extra = ["-ffp-contract=fast" if src in (VMATH, VMATH512) else "-ffp-contract=off"]
            if src in (AVX512, VMATH512) and platform.machine() in ("x86_64", "AMD64"):
                extra += ["-mavx512f", "-mavx512dq", "-mavx512vl", "-mavx2", "-mfma"]
            subprocess.run(["gcc", *flags, *extra, "-c", "-o", str(obj), str(src)], check=True)
            objs.append(obj)
        try:
            subprocess.run(["gcc", "-shared", "-fopenmp", "-o", str(tmp), *map(str, objs), "-lm"], check=True)
        finally:
            for obj in objs:
                obj.unlink(missing_ok=True)
        tmp.replace(LIB)
    return LIB


def _libm_tables():
    """(exp header, exp table, log header, log table) of the libm loaded in this process, read from the library file:
    glibc's ``__exp_data`` / ``__log_data`` (sysdeps/ieee754/dbl-64/e_exp_data.c, e_log_data.c).  They are found by
    their leading constants; ``fm_vmath_init`` then checks the vector routines built on them against libm itself, so
    a layout this does not know simply leaves the vector pass switched off."""
    import struct

    path = None
    try:
        ctypes.CDLL("libm.so.6")
        with open("/proc/self/maps") as f:
            for line in f:
                if "/libm.so" in line or "/libm-" in line:
                    path = line.split()[-1]
                    break
        data = Path(path).read_bytes()
    except (OSError, TypeError):
        return None
    fh = float.fromhex
    e = data.find(struct.pack("<dd", fh("0x1.71547652b82fep+7"), fh("0x1.8p52")))  # N / ln2, shift
    if e < 0:
        return None
    t = data.find(struct.pack("<QQ", 0, 0x3FF0000000000000), e, e + 4096)  # 2^(0/128): tail 0, scale 1.0
    if t < 0:
        return None
    pat = struct.pack("<dd", fh("0x1.62e42fefa3800p-1"), fh("0x1.ef35793c76730p-45"))  # ln2 hi, lo
    pos, found = data.find(pat), -1
    while pos >= 0:
        head = np.frombuffer(data, dtype=np.float64, count=18, offset=pos)
        if head[7] == -0.5 and abs(head[2] + 0.5) < 1e-9 and abs(head[8] - 1 / 3) < 1e-9:  # B[0], A[0], B[1]
            found = pos
            break
        pos = data.find(pat, pos + 8)
    if found < 0 or t + 2048 > len(data) or found + 144 + 2048 > len(data):
        return None
    return (np.frombuffer(data, dtype=np.float64, count=8, offset=e).copy(),
            np.frombuffer(data, dtype=np.uint64, count=256, offset=t).copy(),
            np.frombuffer(data, dtype=np.float64, count=18, offset=found).copy(),
            np.frombuffer(data, dtype=np.float64, count=256, offset=found + 144).copy())


def default_threads() -> int:
    """Threads of the element-wise passes of one fit (``DAJIN2_B200_MLP_THREADS`` overrides).  Three fits run at a time
    (one per mutation type) and, in a multi-GPU run, on rank 0 only, so each gets a third of the cores left after one
    per local rank: 4 on a 16-core box, up to 8 on larger ones.  The thread count cannot change a bit of the result."""
    raw = os.environ.get("DAJIN2_B200_MLP_THREADS")
    if raw:
        return max(1, int(raw))
    cores = os.cpu_count() or 8
    ranks = int(os.environ.get("LOCAL_WORLD_SIZE") or 1)
    return max(2, min(8, (cores - ranks - 2) // 3))


def load_host_lib():
    global _lib
    if _lib is None:
        alt = os.environ.get("DAJIN2_B200_HOSTLIB")  # an alternative build of the same helper (experiments under scripts/)
        if not alt:
            build()
        lib = ctypes.CDLL(alt or str(LIB))
        vp, i64, ci, cd = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_double
        lib.fm_set_threads.argtypes = [ci]
        lib.fm_set_threads.restype = None
        lib.fm_forward_packed.argtypes = [vp, i64, i64, i64, ci, ci, ci, vp, vp, vp]
        lib.fm_forward_packed.restype = ci
        lib.fm_output.argtypes = [vp, cd, vp, i64, vp, vp, vp, ci]
        lib.fm_output.restype = cd
        lib.fm_vmath_init.argtypes = [vp, vp, vp, vp, i64]
        lib.fm_vmath_init.restype = i64
        lib.fm_vmath_ready.argtypes = []
        lib.fm_vmath_ready.restype = ci
        lib.fm512_vmath_init.argtypes = [vp, vp, vp, vp, i64]
        lib.fm512_vmath_init.restype = i64
        lib.fm512_vmath_ready.argtypes = []
        lib.fm512_vmath_ready.restype = ci
        lib.fm_set_avx512.argtypes = [ci]
        lib.fm_set_avx512.restype = None
        lib.fm_get_avx512.argtypes = []
        lib.fm_get_avx512.restype = ci
        lib.fm_set_avx512(int(os.environ.get("DAJIN2_B200_MLP_AVX512", "1") != "0"))
        lib.fm_backward_packed.argtypes = [vp, i64, i64, i64, ci, ci, ci, vp, vp, vp, vp, vp, vp, vp, cd, cd, vp]
        lib.fm_backward_packed.restype = ci
        lib.fm_eval.argtypes = [vp, vp, vp, vp]
        lib.fm_eval.restype = cd
        lib.fm_uniform_choice.argtypes = [i64, vp, ci, vp]
        lib.fm_uniform_choice.restype = ci
        lib.fm_uniform_cumsum_term.argtypes = [i64, i64]
        lib.fm_uniform_cumsum_term.restype = cd
        lib.dj_jsonl_count_lines.argtypes = [vp, i64, ci]
        lib.dj_jsonl_count_lines.restype = i64
        lib.dj_jsonl_index.argtypes = [vp, i64, i64, vp, vp, vp, vp, vp, vp, ci]
        lib.dj_jsonl_index.restype = i64
        lib.dj_gather_fixed.argtypes = [vp, vp, vp, i64, ctypes.c_int32, vp, ci]
        lib.dj_gather_fixed.restype = None
        lib.dj_midsv_postfix.argtypes = [vp, vp, vp, i64, vp, i64, ctypes.c_int32, cd, vp, vp, vp, ci]
        lib.dj_midsv_postfix.restype = i64
        lib.dj_masked_row_sums.argtypes = [vp, i64, i64, ctypes.c_int32, vp, ctypes.c_int32, vp, vp, ci]
        lib.dj_masked_row_sums.restype = None
        lib.dj_sam_count.argtypes = [vp, i64, ci]
        lib.dj_sam_count.restype = i64
        lib.dj_sam_index.argtypes = [vp, i64, i64, vp, vp, vp, vp, vp, vp, vp, vp, ci]
        lib.dj_sam_index.restype = i64
        lib.dj_cs_pack.argtypes = [vp, vp, vp, vp, i64, ci, vp, vp, ci]
        lib.dj_cs_pack.restype = None
        lib.fm_set_threads(default_threads())
        lib.fm_set_forward_threads.argtypes = [ci]
        lib.fm_set_forward_threads.restype = None
        # the AVX-512 forward pass shares its rows among the fit's threads (rows are independent: same bits)
        lib.fm_set_forward_threads(int(os.environ.get("DAJIN2_B200_MLP_FORWARD_THREADS", "1") != "0"))
        if os.environ.get("DAJIN2_B200_MLP_VMATH", "1") != "0":
            tables = _libm_tables()
            if tables is not None:  # 0 differences in the self test switches the four-wide exp / log pass on
                lib.fm_vmath_init(*(t.ctypes.data for t in tables), 250_000)
                if lib.fm_get_avx512():  # and the eight-wide one, on the same terms
                    lib.fm512_vmath_init(*(t.ctypes.data for t in tables), 125_000)
        lib.fm_pairwise_sum.argtypes = [ctypes.c_void_p, ctypes.c_int64]
        lib.fm_pairwise_sum.restype = ctypes.c_double
        _lib = lib
    return _lib


class _EvalCtx(ctypes.Structure):  # FmEvalCtx of hostc/fastmlp.c
    _fields_ = [("X", ctypes.c_void_p), ("xs_row", ctypes.c_int64), ("xs_col", ctypes.c_int64), ("n", ctypes.c_int64),
                ("n_in", ctypes.c_int32), ("h1", ctypes.c_int32), ("h2", ctypes.c_int32), ("vec", ctypes.c_int32),
                ("y", ctypes.c_void_p), ("a1", ctypes.c_void_p), ("a2", ctypes.c_void_p), ("a3", ctypes.c_void_p),
                ("d3", ctypes.c_void_p), ("d2", ctypes.c_void_p), ("d1", ctypes.c_void_p), ("terms", ctypes.c_void_p),
                ("alpha", ctypes.c_double), ("dgemv", ctypes.c_void_p), ("ddot", ctypes.c_void_p)]


_BLAS: tuple | None | bool = False  # False: not looked for yet; None: not usable; (lib, dgemv address, ddot address)


def _numpy_blas():
    """Entry points of the OpenBLAS instance numpy itself is linked against (``cblas_dgemv`` / ``cblas_ddot`` under the
    names its build exports), for ``fm_eval``: the two matrix-vector products and the three dot products of one
    evaluation are then made by the same code numpy would run, with the arguments numpy's ``matmul`` / ``dot`` pass.
    Used only if, on random matrices of the shapes of a fit, both products and the dot return numpy's bits; ``None``
    otherwise (the evaluation then goes through numpy as before).  ``DAJIN2_B200_MLP_ONECALL=0`` switches it off."""
    global _BLAS
    if _BLAS is not False:
        return _BLAS
    _BLAS = None
    if os.environ.get("DAJIN2_B200_MLP_ONECALL", "1") == "0":
        return None
    try:
        from threadpoolctl import threadpool_info

        paths = [i["filepath"] for i in threadpool_info() if i.get("user_api") == "blas" and "numpy" in i["filepath"]]
        if len(paths) != 1:
            return None
        lib = ctypes.CDLL(paths[0])
        i64, dbl, vp, ci = ctypes.c_int64, ctypes.c_double, ctypes.c_void_p, ctypes.c_int
        for prefix, suffix in (("scipy_", "64_"), ("", "64_"), ("scipy_", ""), ("", "")):
            if hasattr(lib, f"{prefix}cblas_dgemv{suffix}") and hasattr(lib, f"{prefix}cblas_ddot{suffix}"):
                break
        else:
            return None
        if suffix != "64_":  # fm_eval passes 64-bit integers (numpy's wheels are built against the ILP64 interface)
            return None
        dgemv, ddot = getattr(lib, f"{prefix}cblas_dgemv{suffix}"), getattr(lib, f"{prefix}cblas_ddot{suffix}")
        dgemv.argtypes = [ci, ci, i64, i64, dbl, vp, i64, vp, i64, dbl, vp, i64]
        dgemv.restype = None
        ddot.argtypes = [i64, vp, i64, vp, i64]
        ddot.restype = dbl
        rng = np.random.default_rng(11)
        for trial in range(24):
            n, h = int(rng.integers(500, 40_000)), int(rng.integers(2, 7))  # h = 1 takes other routes inside numpy
            a2 = np.maximum(rng.normal(size=(n, h)), 0) * rng.uniform(0.01, 60)
            packed = rng.normal(size=64)
            w3 = packed[7:7 + h].reshape(h, 1)
            d3 = rng.normal(size=(n, 1)) * 10.0 ** rng.integers(-6, 1)
            y1, y2 = np.empty((n, 1)), np.empty((h, 1))
            dgemv(102, 112, h, n, 1.0, a2.ctypes.data, h, w3.ctypes.data, 1, 0.0, y1.ctypes.data, 1)
            dgemv(101, 112, n, h, 1.0, a2.ctypes.data, h, d3.ctypes.data, 1, 0.0, y2.ctypes.data, 1)
            sl = packed[3:3 + int(rng.integers(1, 30))]
            if not (np.array_equal((a2 @ w3).view(np.uint64), y1.view(np.uint64))
                    and np.array_equal((a2.T @ d3).view(np.uint64), y2.view(np.uint64))
                    and float(np.dot(sl, sl)) == ddot(sl.shape[0], sl.ctypes.data, 1, sl.ctypes.data, 1)):
                return None
        _BLAS = (lib, ctypes.cast(dgemv, vp).value, ctypes.cast(ddot, vp).value)
    except Exception:  # noqa: BLE001 - anything unexpected: the evaluation keeps going through numpy
        _BLAS = None
    return _BLAS


def _ptr(a: np.ndarray) -> int:
    return a.__array_interface__["data"][0]  # the address, without building the ctypes helper object each time


_BINARIZERS: dict = {}  # label dtype -> LabelBinarizer fitted on [0, 1]


class FastMLPClassifier(MLPClassifier):
    """``MLPClassifier`` whose lbfgs loss/gradient evaluation is fused (two hidden relu layers, one logistic output,
    float64, no sample weights); any other configuration uses scikit-learn's evaluation."""

    fast_path_used_: bool = False
    fast_validation_used_: bool = False
    #: True: the two comparisons with scikit-learn's own evaluation (first and last evaluated point) are NOT made inside
    #: the fit; the points and the fused results are left in ``deferred_checks_`` for the caller, who has them evaluated
    #: by scikit-learn elsewhere (``mlp_pool``: another worker process, beside the fit) and compares the bits
    defer_checks: bool = False
    deferred_checks_: list | None = None

    def _validate_input(self, X, y, incremental, reset):
        """scikit-learn's input validation spends several ms per fit in the generic label machinery (sparse label
        binarisation of 2 x (10 000 + L) labels).  For the one shape the reference feeds it -- a finite float64
        matrix in either memory order (the reference's concatenation of ``.T`` views is column-major) and 1-D integer labels that are exactly {0, 1} -- the same attributes and the same (n, 1) boolean target
        are produced directly; anything else goes through scikit-learn's own method."""
        from sklearn.preprocessing import LabelBinarizer

        if (not incremental and reset and not self.warm_start and isinstance(X, np.ndarray) and X.ndim == 2
                and X.dtype == np.float64 and (X.flags.c_contiguous or X.flags.f_contiguous) and isinstance(y, np.ndarray) and y.ndim == 1
                and y.dtype.kind == "i" and y.shape[0] == X.shape[0] and X.shape[0] > 1):
            ones = y == 1
            if (ones | (y == 0)).all() and ones.any() and not ones.all() and np.isfinite(X).all():
                self.n_features_in_ = X.shape[1]
                fitted = _BINARIZERS.get(y.dtype)
                if fitted is None:  # the fit of a binariser on [0, 1] costs 1.6 ms: once per label dtype and process
                    fitted = _BINARIZERS.setdefault(y.dtype, LabelBinarizer().fit(np.array([0, 1], dtype=y.dtype)))
                self._label_binarizer = copy.copy(fitted)
                self._label_binarizer.classes_ = fitted.classes_.copy()
                self.fast_validation_used_ = True
                self.classes_ = self._label_binarizer.classes_
                return X, ones.reshape(-1, 1)
        return super()._validate_input(X, y, incremental, reset)

    def _fast_applicable(self, X, y, sample_weight) -> bool:
        return (sample_weight is None and self.activation == "relu" and self.out_activation_ == "logistic"
                and self.n_layers_ == 4 and self.n_outputs_ == 1 and isinstance(X, np.ndarray)
                and X.dtype == np.float64 and X.ndim == 2 and self.loss == "log_loss")

    def _fast_setup(self, X, y):
        n, n_in = X.shape
        h1, h2 = self.coefs_[0].shape[1], self.coefs_[1].shape[1]
        f = lambda *shape: np.empty(shape, dtype=np.float64)  # noqa: E731
        fm = {
            "n": n, "lib": load_host_lib(), "dims": (n_in, h1, h2), "y": np.ascontiguousarray(y, dtype=np.float64).reshape(n, 1),
            "a1": f(n, h1), "a2": f(n, h2), "d3": f(n, 1), "d2": f(n, h2), "d1": f(n, h1), "terms": f(n), "sum3": f(1),
            "xs": (X.strides[0] // 8, X.strides[1] // 8), "X": X,
            # offsets of W1, W2, W3, b1, b2, b3 in scikit-learn's packed parameter vector (_pack / _unpack)
            "o_w3": n_in * h1 + h1 * h2, "o_b3": n_in * h1 + h1 * h2 + h2 + h1 + h2,
            "n_par": n_in * h1 + h1 * h2 + h2 + h1 + h2 + 1,
        }
        yy = fm["y"]
        fm["vec"] = 0
        if bool(((yy == 0.0) | (yy == 1.0)).all()):  # 2: eight rows at a time (AVX-512), 1: four, 0: libm's own calls
            lib = fm["lib"]
            fm["vec"] = 2 if (lib.fm_get_avx512() and lib.fm512_vmath_ready()) else int(bool(lib.fm_vmath_ready()))
        fm["a2T"] = fm["a2"].T
        fm["ptr"] = {k: _ptr(fm[k]) for k in ("a1", "a2", "d3", "d2", "d1", "terms", "sum3", "y", "X")}
        fm["ctx"] = None
        blas = _numpy_blas()
        if blas is not None and h2 >= 2:  # one C call per evaluation (fm_eval), the products by numpy's own OpenBLAS
            fm["a3"] = f(n, 1)
            ptr = fm["ptr"]
            fm["ctx"] = _EvalCtx(ptr["X"], X.strides[0] // 8, X.strides[1] // 8, n, n_in, h1, h2, fm["vec"], ptr["y"],
                                 ptr["a1"], ptr["a2"], _ptr(fm["a3"]), ptr["d3"], ptr["d2"], ptr["d1"], ptr["terms"],
                                 float(self.alpha), blas[1], blas[2])
            fm["ctx_ref"] = ctypes.addressof(fm["ctx"])
            fm["status"] = ctypes.c_int(0)
            fm["status_ref"] = ctypes.addressof(fm["status"])
        self._fm = fm

    def _fast_loss_grad(self, packed, X):
        fm = self._fm
        lib, n, ptr = fm["lib"], fm["n"], fm["ptr"]
        n_in, h1, h2 = fm["dims"]
        xs_row, xs_col = fm["xs"]
        if X is not fm["X"] or packed.dtype != np.float64 or not packed.flags.c_contiguous or packed.shape[0] != fm["n_par"]:
            raise RuntimeError("unexpected arguments for the fused evaluation")
        pp = _ptr(packed)
        if fm["ctx"] is not None:
            if fm["ctx"].vec != fm["vec"]:
                fm["ctx"].vec = fm["vec"]
            grad = np.empty(fm["n_par"], dtype=np.float64)
            loss = np.float64(lib.fm_eval(fm["ctx_ref"], pp, _ptr(grad), fm["status_ref"]))
            if fm["status"].value != 0:
                raise RuntimeError("layer too wide for the fused evaluation")
            return loss, grad
        o_w3, o_b3 = fm["o_w3"], fm["o_b3"]
        if lib.fm_forward_packed(ptr["X"], xs_row, xs_col, n, n_in, h1, h2, pp, ptr["a1"], ptr["a2"]) != 0:
            raise RuntimeError("layer too wide for the fused evaluation")
        W3 = packed[o_w3:o_w3 + h2].reshape(h2, 1)  # the view _unpack makes of the output unit's weights
        a3 = fm["a2"] @ W3  # (n, h2) @ (h2, 1): numpy -> dgemv, as safe_sparse_dot does
        loss = np.float64(lib.fm_output(_ptr(a3), float(packed[o_b3]), ptr["y"], n, ptr["d3"], ptr["terms"], ptr["sum3"],
                                      fm["vec"]))
        values = 0
        lo = 0
        for size in (n_in * h1, h1 * h2, h2):  # for s in self.coefs_: values += np.dot(s.ravel(), s.ravel())
            sl = packed[lo:lo + size]
            values += np.dot(sl, sl)
            lo += size
        loss += (0.5 * self.alpha) * values / n
        g3_raw = fm["a2T"] @ fm["d3"]  # (h2, n) @ (n, 1): numpy -> dgemv
        grad = np.empty(fm["n_par"], dtype=np.float64)
        if lib.fm_backward_packed(ptr["X"], xs_row, xs_col, n, n_in, h1, h2, pp, ptr["a1"], ptr["a2"], ptr["d3"], ptr["d2"],
                                  ptr["d1"], _ptr(g3_raw), float(fm["sum3"][0]), float(self.alpha), _ptr(grad)) != 0:
            raise RuntimeError("layer too wide for the fused evaluation")
        return loss, grad

    def _loss_grad_lbfgs(self, packed_coef_inter, X, y, sample_weight, activations, deltas, coef_grads, intercept_grads):
        mode = getattr(self, "_fm_mode", None)
        if mode is None:  # first evaluation of this fit: decide, and check against scikit-learn's own evaluation
            mode = "sklearn"
            if self._fast_applicable(X, y, sample_weight) and self.defer_checks:
                try:
                    self._fast_setup(X, y)
                    loss, grad = self._fast_loss_grad(packed_coef_inter, X)
                    self.deferred_checks_ = [(packed_coef_inter.copy(), float(loss), grad.copy())]
                    mode = "fast"
                except Exception:  # noqa: BLE001
                    mode = "sklearn"
            elif self._fast_applicable(X, y, sample_weight):
                self._fast_setup(X, y)
                ref = super()._loss_grad_lbfgs(packed_coef_inter.copy(), X, y, sample_weight, activations, deltas,
                                               coef_grads, intercept_grads)
                try:
                    got = self._fast_loss_grad(packed_coef_inter, X)
                    if not _same_bits(ref, got) and self._fm["vec"]:
                        self._fm["vec"] = 0  # the four-wide exp / log pass is not libm's here: libm's own calls
                        got = self._fast_loss_grad(packed_coef_inter, X)
                    if _same_bits(ref, got):
                        mode = "fast"
                except Exception:  # noqa: BLE001 - anything unexpected: scikit-learn's own evaluation is always right
                    mode = "sklearn"
            self._fm_mode = mode
        if mode == "fast":
            self._fm_last = packed_coef_inter.copy()
            return self._fast_loss_grad(packed_coef_inter, X)
        return super()._loss_grad_lbfgs(packed_coef_inter, X, y, sample_weight, activations, deltas, coef_grads,
                                        intercept_grads)

    def _fit_lbfgs(self, X, y, sample_weight, activations, deltas, coef_grads, intercept_grads, layer_units):
        self._fm_mode = None
        start = [c.copy() for c in self.coefs_], [b.copy() for b in self.intercepts_]
        self.lean_driver_used_ = False
        if os.environ.get("DAJIN2_B200_MLP_LEAN", "1") != "0" and _lean_selftest():
            # scikit-learn's _fit_lbfgs looks `scipy.optimize.minimize` up at call time: for the duration of this fit it
            # finds the lean driver (same compiled L-BFGS-B, same loop; self-tested against scipy's above)
            import scipy.optimize

            stock = scipy.optimize.minimize

            def lean(fun, x0, method=None, jac=None, options=None, args=(), **kw):
                opts = dict(options or {})
                if method != "L-BFGS-B" or jac is not True or kw or set(opts) - {"maxfun", "maxiter", "gtol", "iprint"} \
                        or opts.get("iprint", -1) not in (-1, None):
                    return stock(fun, x0, method=method, jac=jac, options=options, args=args, **kw)
                self.lean_driver_used_ = True
                return _lean_lbfgsb(fun, x0, args, opts.get("maxfun", 15000), opts.get("maxiter", 15000), opts.get("gtol", 1e-5))

            scipy.optimize.minimize = lean
            try:
                super()._fit_lbfgs(X, y, sample_weight, activations, deltas, coef_grads, intercept_grads, layer_units)
            finally:
                scipy.optimize.minimize = stock
        else:
            super()._fit_lbfgs(X, y, sample_weight, activations, deltas, coef_grads, intercept_grads, layer_units)
        self.fast_path_used_ = self._fm_mode == "fast"
        if self._fm_mode == "fast" and self.defer_checks:
            last = self._fm_last
            final = [c.copy() for c in self.coefs_], [b.copy() for b in self.intercepts_]
            loss, grad = self._fast_loss_grad(last, X)
            self.deferred_checks_.append((last.copy(), float(loss), grad.copy()))
            self.coefs_, self.intercepts_ = final
        elif self._fm_mode == "fast":
            # check the last evaluated point as well (late iterates exercise the clipped / saturated branches)
            last = self._fm_last
            final = [c.copy() for c in self.coefs_], [b.copy() for b in self.intercepts_]
            ref = super()._loss_grad_lbfgs(last.copy(), X, y, sample_weight, activations, deltas, coef_grads,
                                           intercept_grads)
            got = self._fast_loss_grad(last, X)
            if _same_bits(ref, got):
                self.coefs_, self.intercepts_ = final
            else:  # not the arithmetic this was written against: redo the fit with scikit-learn's evaluation
                self.coefs_, self.intercepts_ = start
                self._fm_mode = "sklearn"
                self.fast_path_used_ = False
                MLPClassifier._fit_lbfgs(self, X, y, sample_weight, activations, deltas, coef_grads, intercept_grads,
                                         layer_units)
        self._fm = None


# ---------------------------------------------------------------------------------------------------------
# scipy's L-BFGS-B driver loop without its per-evaluation Python machinery
# ---------------------------------------------------------------------------------------------------------
_LEAN_OK: bool | None = None  # None: not self-tested yet


def _lean_lbfgsb(fun, x0, args, maxfun, maxiter, gtol, maxcor=10, ftol=2.220446049250313e-09, maxls=20):
    """``scipy.optimize.minimize(fun, x0, method="L-BFGS-B", jac=True, args=args, options={maxfun, maxiter, gtol})``
    with the SAME compiled routine (``scipy.optimize._lbfgsb.setulb``), the same workspace and the same driver loop
    as ``scipy/optimize/_lbfgsb_py.py:_minimize_lbfgsb`` -- minus ``ScalarFunction`` / ``MemoizeJac``, which spend
    ~0.1 ms of pure Python per evaluation (a fifth of a fit of the reference's MLP).  The iterates only depend on
    what ``setulb`` is fed, so they are scipy's; ``_lean_selftest`` compares whole runs bit for bit before this is used."""
    from scipy.optimize import OptimizeResult, _lbfgsb
    from scipy.optimize._lbfgsb_py import status_messages, task_messages

    m = maxcor
    factr = ftol / np.finfo(float).eps
    x0 = np.asarray(x0).ravel()
    (n,) = x0.shape
    int_dtype = _lbfgsb_int_dtype()
    nbd = np.zeros(n, dtype=int_dtype)
    low_bnd = np.zeros(n, dtype=np.float64)
    upper_bnd = np.zeros(n, dtype=np.float64)
    x = np.array(x0, dtype=np.float64)
    f = np.array(0.0, dtype=np.float64)
    g = np.zeros((n,), dtype=np.float64)
    wa = np.zeros(2 * m * n + 5 * n + 11 * m * m + 8 * m, np.float64)
    iwa = np.zeros(3 * n, dtype=int_dtype)
    task = np.zeros(2, dtype=int_dtype)
    ln_task = np.zeros(2, dtype=int_dtype)
    lsave = np.zeros(4, dtype=int_dtype)
    isave = np.zeros(44, dtype=int_dtype)
    dsave = np.zeros(29, dtype=np.float64)
    n_iterations = 0
    nfev = 0
    while True:
        g = g.astype(np.float64)
        _lbfgsb.setulb(m, x, low_bnd, upper_bnd, nbd, f, g, factr, gtol, wa, iwa, task, lsave, isave, dsave, maxls, ln_task)
        if task[0] == 3:
            # ScalarFunction hands the objective a copy of x and keeps f as a Python float, g as a float64 array
            f, g = fun(np.copy(x), *args)
            nfev += 1
            f = float(f)
            g = np.atleast_1d(np.asarray(g, dtype=np.float64))
        elif task[0] == 1:
            n_iterations += 1
            if n_iterations >= maxiter:
                task[0] = 5
                task[1] = 504
            elif nfev > maxfun:
                task[0] = 5
                task[1] = 502
        else:
            break
    if task[0] == 4:
        warnflag = 0
    elif nfev > maxfun or n_iterations >= maxiter:
        warnflag = 1
    else:
        warnflag = 2
    msg = status_messages[task[0]] + ": " + task_messages[task[1]]
    return OptimizeResult(fun=f, jac=g, nfev=nfev, njev=nfev, nit=n_iterations, status=warnflag, message=msg, x=x,
                          success=(warnflag == 0))


def _lbfgsb_int_dtype():
    from scipy.optimize import _lbfgsb_py

    return np.int64 if getattr(_lbfgsb_py, "HAS_ILP64", False) else np.int32


def _lean_selftest() -> bool:
    """Whole runs of the lean driver against ``scipy.optimize.minimize`` on three small problems: x, f, g, nit, nfev,
    status and message must agree exactly.  Any exception (another scipy with another ``setulb``) switches it off."""
    global _LEAN_OK
    if _LEAN_OK is not None:
        return _LEAN_OK
    try:
        import scipy.optimize

        rng = np.random.default_rng(7)
        ok = True
        for n, scale, maxiter in ((27, 1.0, 1000), (9, 30.0, 1000), (27, 5.0, 7)):
            A = rng.normal(size=(n, n))
            A = A @ A.T / n + np.eye(n) * 0.1
            b = rng.normal(size=n) * scale

            def fun(x, A=A, b=b):
                r = A @ x - b
                q = np.tanh(x)
                return 0.5 * float(r @ r) + float(np.sum(np.log1p(q * q))), A.T @ r + 2 * q * (1 - q * q) / (1 + q * q)

            x0 = rng.normal(size=n)
            ref = scipy.optimize.minimize(fun, x0.copy(), method="L-BFGS-B", jac=True,
                                          options={"maxfun": 15000, "maxiter": maxiter, "gtol": 1e-4})
            got = _lean_lbfgsb(fun, x0.copy(), (), 15000, maxiter, 1e-4)
            ok = ok and (np.array_equal(ref.x.view(np.uint64), got.x.view(np.uint64)) and float(ref.fun) == float(got.fun)
                         and np.array_equal(np.asarray(ref.jac).view(np.uint64), np.asarray(got.jac).view(np.uint64))
                         and ref.nit == got.nit and ref.nfev == got.nfev and ref.status == got.status
                         and ref.message == got.message)
        _LEAN_OK = bool(ok)
    except Exception:  # noqa: BLE001
        _LEAN_OK = False
    return _LEAN_OK


def _same_bits(ref, got) -> bool:
    (l0, g0), (l1, g1) = ref, got
    a = np.asarray(l0, dtype=np.float64).reshape(1).view(np.uint64)
    b = np.asarray(l1, dtype=np.float64).reshape(1).view(np.uint64)
    return bool(a[0] == b[0]) and g0.shape == g1.shape and bool(
        np.array_equal(np.ascontiguousarray(g0, dtype=np.float64).view(np.uint64),
                       np.ascontiguousarray(g1, dtype=np.float64).view(np.uint64)))


class ProbeMLPClassifier(MLPClassifier):
    """scikit-learn's own loss / gradient (``_loss_grad_lbfgs``) at given parameter vectors of the model a fit on
    (X, y) would start from: ``fit`` runs scikit-learn's validation, initialisation and buffer set-up, then evaluates
    instead of optimising.  ``points_`` = [(packed, loss, grad), ...]; ``None`` among ``points`` stands for the
    initial parameters (random_state decides them, so a fit elsewhere starts from the same ones)."""

    def __init__(self, points=(None,), **kw):
        super().__init__(**kw)
        self.points = points

    def _fit_lbfgs(self, X, y, sample_weight, activations, deltas, coef_grads, intercept_grads, layer_units):
        from sklearn.neural_network._multilayer_perceptron import _pack

        self._coef_indptr, self._intercept_indptr = [], []
        start = 0
        for i in range(self.n_layers_ - 1):  # as MLPClassifier._fit_lbfgs lays the packed vector out
            n_fan_in, n_fan_out = layer_units[i], layer_units[i + 1]
            end = start + (n_fan_in * n_fan_out)
            self._coef_indptr.append((start, end, (n_fan_in, n_fan_out)))
            start = end
        for i in range(self.n_layers_ - 1):
            end = start + layer_units[i + 1]
            self._intercept_indptr.append((start, end))
            start = end
        x0 = _pack(self.coefs_, self.intercepts_)
        self.points_ = []
        for p in self.points:
            packed = x0.copy() if p is None else np.array(p, dtype=np.float64)
            loss, grad = MLPClassifier._loss_grad_lbfgs(self, packed.copy(), X, y, sample_weight, activations, deltas,
                                                        coef_grads, intercept_grads)
            self.points_.append((packed, float(loss), np.array(grad, dtype=np.float64)))
        self._unpack(x0)
        self.n_iter_ = 0
        self.loss_ = float("nan")
