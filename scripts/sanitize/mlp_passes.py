import os, sys
sys.path.insert(0, str(__import__('pathlib').Path(__file__).resolve().parents[2]))
import numpy as np
from dajin2_b200 import fastmlp
from pathlib import Path
fastmlp.LIB = Path("/tmp/libdjhost_asan.so")
fastmlp.build = lambda force=False: fastmlp.LIB
lib = fastmlp.load_host_lib(); P = fastmlp._ptr
print("avx512", lib.fm_get_avx512(), "v4", lib.fm_vmath_ready(), "v8", lib.fm512_vmath_ready())
rng = np.random.default_rng(0)
for n in (1, 5, 7, 8, 9, 15, 16, 17, 127, 128, 129, 1000, 1003):
    for order in ("C", "F"):
        for on in (0, 1):
            lib.fm_set_avx512(on)
            X = np.array(rng.uniform(0, 100, (n, 2)), order=order); packed = rng.normal(0, 0.7, 30); d3 = rng.normal(0, 1, (n, 1))
            a1 = np.empty((n, 5)); a2 = np.empty((n, 2)); d2 = np.empty((n, 2)); d1 = np.empty((n, 5)); grad = np.empty(30)
            assert lib.fm_forward_packed(P(X), X.strides[0]//8, X.strides[1]//8, n, 2, 5, 2, P(packed), P(a1), P(a2)) == 0
            g3 = a2.T @ d3
            assert lib.fm_backward_packed(P(X), X.strides[0]//8, X.strides[1]//8, n, 2, 5, 2, P(packed), P(a1), P(a2), P(d3), P(d2), P(d1), P(g3), 0.5, 1e-5, P(grad)) == 0
            z = rng.normal(0, 8, (n, 1)); y = (rng.random((n, 1)) < 0.5).astype(np.float64); d = np.empty((n, 1)); t = np.empty(n); s = np.empty(1)
            for vec in (0, 1, 2):
                zz = z.copy(); lib.fm_output(P(zz), 0.1, P(y), n, P(d), P(t), P(s), vec)
print("asan mlp passes complete")
