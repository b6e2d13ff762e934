#!/usr/bin/env python
"""Predict throughput of the streaming SIMT kernel for the sweep's large reservoirs (cfg4: 1024 and 2048
neurons, 4x8, T = 522), fp32, device-resident inputs.  TFLOP/s = algorithmic (2 N (N + n_in +
n_out) + 2 n_out (N + n_in) per step and frame)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "esn-ofdm-mimo_b200"))
from esn_b200 import Reservoir  # noqa: E402

ni, no, T = 16, 8, 522
for N in (512, 1024, 2048):
    rng = np.random.RandomState(0)
    W = (rng.rand(N, N) - 0.5) * (0.9 * 2 / np.sqrt(N / 3))          # ~ spectral radius 0.9 without the eigen solve
    res = Reservoir(W, rng.rand(N, ni) * 2 - 1, rng.rand(N, no) * 2 - 1, 0.005 * np.ones(ni), np.zeros(ni),
                    5e-7 * np.ones(no), np.zeros(no), 0.001, True)
    Wo = torch.randn(1, no, N + ni, device="cuda") * 1e-6
    for B in (592, 2368, 9472):
        us = torch.randn(B, T, ni, device="cuda")
        res.predict(us, Wo, transient=10, precision="fp32", seed=3)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res.predict(us, Wo, transient=10, precision="fp32", seed=3)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        flop = B * T * (2 * N * (N + ni + no) + 2 * no * (N + ni))
        print(f"N={N:5d} B={B:5d}: {ms:9.2f} ms  {B / ms * 1e3:10.0f} sym/s  {flop / ms / 1e9:7.2f} TFLOP/s")
