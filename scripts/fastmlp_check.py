#!/usr/bin/env python
"""FastMLPClassifier against scikit-learn's MLPClassifier on this machine: identical parameters? how much faster?"""
import os, sys, time, warnings
for v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
    os.environ.setdefault(v, "1")
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent)); sys.path.insert(0, str(Path(__file__).resolve().parent))
import numpy as np
warnings.simplefilter("ignore")
from threadpoolctl import threadpool_info
from sklearn.neural_network import MLPClassifier
from mlp_case import case
from dajin2_b200 import fastmlp
print([(d.get("internal_api"), d.get("architecture"), d.get("version")) for d in threadpool_info()], "cpus", os.cpu_count())
os.system("grep -m1 'model name' /proc/cpuinfo")
fastmlp.build(force=True)
kw = dict(solver="lbfgs", alpha=1e-5, hidden_layer_sizes=(5, 2), random_state=1, max_iter=1000)
for L, seed in ((5000, 5), (5000, 6), (1000, 7), (2832, 8), (300, 9), (4309, 10)):
    X, y = case(L, seed)
    t = time.perf_counter(); ref = MLPClassifier(**kw).fit(X, y); t_ref = time.perf_counter() - t
    t = time.perf_counter(); m = fastmlp.FastMLPClassifier(**kw).fit(X, y); t_fast = time.perf_counter() - t
    same = all(np.array_equal(a, b) for a, b in zip(ref.coefs_ + ref.intercepts_, m.coefs_ + m.intercepts_))
    print(L, "sklearn %.3f s  fast %.3f s  n_iter %d/%d fast_path %s identical %s" % (t_ref, t_fast, ref.n_iter_, m.n_iter_, m.fast_path_used_, same), flush=True)
