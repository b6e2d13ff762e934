#!/usr/bin/env python
"""cfg2 (BASELINE.json configs[1]): 2x2 16-QAM, 100-neuron reservoir, 4096 frames per step, T = 522.
Times predict on the tensor-core path and on the fp32 SIMT path.  python profiles/cfg2_bench.py [N_res]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "esn-ofdm-mimo_b200"))
from esn_b200 import Reservoir  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 100
ni, no, T, B = 4, 4, 522, 4096
rng = np.random.RandomState(42)
W = rng.rand(N, N) - 0.5
W *= 0.9 / np.max(np.abs(np.linalg.eigvals(W)))
res = Reservoir(W, rng.rand(N, ni) * 2 - 1, rng.rand(N, no) * 2 - 1, input_scaling=0.005 * np.ones(ni),
                teacher_scaling=5e-7 * np.ones(no), noise=0.001)
x = torch.randn(B, T, ni, device="cuda")
Wo = torch.randn(1, no, N + ni, dtype=torch.float64, device="cuda") * 1e-6
rd = res.tc_prepare(Wo, res.input_scale_exponent(x))
flop = T * (2 * N * (N + ni + no) + 2 * no * (N + ni))


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for name, fn in (("tc", lambda: res.predict_tc(x, rd, transient=10, seed=1)),
                 ("fp32 simt", lambda: res.predict(x, Wo, transient=10, precision="fp32", seed=1))):
    ms = timed(fn)
    print(f"N={N} {name:10s}: {ms:8.3f} ms per {B} frames -> {B / ms * 1e3:12.0f} OFDM symbols/s, {flop * B / ms / 1e9:8.2f} TFLOP/s algorithmic")
