#!/usr/bin/env python
"""Forward / backward passes of the 2-5-2 MLP evaluation, AVX2 against AVX-512 forms (hostc/fastmlp.c, fastmlp512.c): ns per row at several row counts (row-major X)."""
import os, sys, time
sys.path.insert(0, str(__import__("pathlib").Path(__file__).resolve().parents[2]))
import numpy as np
from dajin2_b200 import fastmlp
lib = fastmlp.load_host_lib(); P = fastmlp._ptr
rng = np.random.default_rng(0)
for n in (2000, 8000, 30000, 120000):
    X = rng.uniform(0, 100, (n, 2)); packed = rng.normal(0, 0.7, 30); d3 = rng.normal(0, 1, (n, 1))
    a1 = np.empty((n, 5)); a2 = np.empty((n, 2)); d2 = np.empty((n, 2)); d1 = np.empty((n, 5)); g3 = np.zeros(2); grad = np.empty(30)
    for on in (0, 1):
        lib.fm_set_avx512(on)
        tf = tb = 1e9
        for _ in range(200):
            t = time.perf_counter(); lib.fm_forward_packed(P(X), 2, 1, n, 2, 5, 2, P(packed), P(a1), P(a2)); tf = min(tf, time.perf_counter() - t)
            t = time.perf_counter(); lib.fm_backward_packed(P(X), 2, 1, n, 2, 5, 2, P(packed), P(a1), P(a2), P(d3), P(d2), P(d1), P(g3), 0.5, 1e-5, P(grad)); tb = min(tb, time.perf_counter() - t)
        print(n, "avx512" if on else "avx2  ", f"forward {tf/n*1e9:.2f} ns/row  backward {tb/n*1e9:.2f} ns/row")
