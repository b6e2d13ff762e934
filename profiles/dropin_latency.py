#!/usr/bin/env python
"""Latency of the unmodified reference call shapes through the drop-in modules (one frame per call, as the
demos do): ESN.fit, ESN.predict and trainMIMOESN_generic at the 4x8 / 512-neuron shape, fp64 (default)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "esn-ofdm-mimo_b200"), os.path.join(ROOT, "esn-ofdm-mimo_b200", "libs")):
    sys.path.insert(0, p)
from pyESN import ESN  # noqa: E402
from helper_mimo_esn_generic import trainMIMOESN_generic  # noqa: E402

N, N_t, N_r, n_res, cp = 512, 4, 8, 512, 7
rng = np.random.RandomState(0)
esn = ESN(n_inputs=2 * N_r, n_outputs=2 * N_t, n_reservoir=n_res, spectral_radius=0.9, sparsity=0.1,
          input_scaling=0.005 * np.ones(2 * N_r), input_shift=np.zeros(2 * N_r),
          teacher_scaling=5e-7 * np.ones(2 * N_t), teacher_shift=np.zeros(2 * N_t), random_state=42)
y_CP = (rng.randn(N + cp, N_r) + 1j * rng.randn(N + cp, N_r))
x_CP = (rng.randn(N + cp, N_t) + 1j * rng.randn(N + cp, N_t))


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


res = trainMIMOESN_generic(esn, 0, 0, 6, cp, N, N_t, N_r, 8, y_CP, x_CP)
ein, eout, nf = res[0], res[1], res[7]
print("trainMIMOESN_generic (2 fits + 1 predict): %.1f ms" % timed(lambda: trainMIMOESN_generic(esn, 0, 0, 6, cp, N, N_t, N_r, 8, y_CP, x_CP)))
print("ESN.fit      [522 x 16]: %.1f ms" % timed(lambda: esn.fit(ein, eout, nf)))
print("ESN.predict  [522 x 16]: %.1f ms" % timed(lambda: esn.predict(ein, nf, continuation=False)))
print("trainMIMOESN_generic with delay scan (7 candidates, batched): %.1f ms" % timed(lambda: trainMIMOESN_generic(esn, 1, 0, 6, cp, N, N_t, N_r, 8, y_CP, x_CP), reps=3))
