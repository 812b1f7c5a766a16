#!/usr/bin/env python
"""The loci-selection stage as bench.py runs it: three fits side by side in the worker pool, against one at a time."""
import os, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from dajin2_b200 import mlp_pool

L = 5000
rng = np.random.default_rng(3)
cases = []
for rate in (0.4, 0.8, 0.5):  # '+', '-', '*' percentages of the r10 profile
    control = np.abs(rng.normal(rate, rate * 0.15, L))
    sample = control + rng.normal(0, rate * 0.05, L)
    sample[[1000, 1800, 2400, 3000, 3500]] += [25, 15, 10, 5, 2]
    cases.append((np.clip(sample, 0, 100), np.minimum(control, np.clip(sample, 0, 100)), 0.1))
mlp_pool.warm_up()
for rep in range(2):
    t0 = time.perf_counter()
    futs = [mlp_pool.submit(*c) for c in cases]
    done = []
    for f in futs:
        f.result(); done.append((time.perf_counter() - t0) * 1e3)
    print("three at a time: finished after", [round(x, 1) for x in done], "ms", flush=True)
    one = []
    for c in cases:
        t = time.perf_counter(); mlp_pool.submit(*c).result(); one.append((time.perf_counter() - t) * 1e3)
    print("one at a time:  ", [round(x, 1) for x in one], "ms", flush=True)
