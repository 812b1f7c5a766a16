#!/usr/bin/env python
"""Where one FastMLPClassifier fit spends its time on this machine (C passes, numpy products, everything else)."""
import os, sys, time, warnings
for v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
    os.environ.setdefault(v, "1")
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent)); sys.path.insert(0, str(Path(__file__).resolve().parent))
warnings.simplefilter("ignore")
import numpy as np
from mlp_case import case
from dajin2_b200 import fastmlp

X, y = case(5000, 5)
kw = dict(solver="lbfgs", alpha=1e-5, hidden_layer_sizes=(5, 2), random_state=1, max_iter=1000)
lib = fastmlp.load_host_lib()
acc = {"forward": 0.0, "output": 0.0, "backward": 0.0, "evals": 0, "loss_grad": 0.0}
for name, key in (("fm_forward_packed", "forward"), ("fm_output", "output"), ("fm_backward_packed", "backward")):
    real = getattr(lib, name)
    def wrap(*a, _real=real, _key=key):
        t = time.perf_counter(); r = _real(*a); acc[_key] += time.perf_counter() - t; return r
    setattr(lib, name, wrap)
orig = fastmlp.FastMLPClassifier._fast_loss_grad
def timed(self, packed, X):
    t = time.perf_counter(); r = orig(self, packed, X); acc["loss_grad"] += time.perf_counter() - t; acc["evals"] += 1; return r
fastmlp.FastMLPClassifier._fast_loss_grad = timed
for thr in (int(a) for a in (sys.argv[1:] or ["4"])):
    lib.fm_set_threads(thr)
    fastmlp.FastMLPClassifier(**kw).fit(X, y)
    for k in acc: acc[k] = 0 if k == "evals" else 0.0
    t = time.perf_counter(); m = fastmlp.FastMLPClassifier(**kw).fit(X, y); total = time.perf_counter() - t
    n = acc["evals"]
    print(f"threads {thr}: fit {total*1e3:.1f} ms, {n} evaluations, n_iter {m.n_iter_}; per evaluation: C forward "
          f"{acc['forward']/n*1e3:.3f} output {acc['output']/n*1e3:.3f} backward {acc['backward']/n*1e3:.3f}, whole "
          f"loss/grad {acc['loss_grad']/n*1e3:.3f} ms; outside loss/grad {(total-acc['loss_grad'])*1e3:.1f} ms", flush=True)
