#!/usr/bin/env python
"""One predict launch of the small-batch cluster kernel (for ncu): python cluster_one.py B precision"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "esn-ofdm-mimo_b200"))
from esn_b200 import Reservoir  # noqa: E402

B, prec = int(sys.argv[1]), sys.argv[2]
N, ni, no, T = 512, 16, 8, 522
dt = torch.float64 if prec == "fp64" else torch.float32
rng = np.random.RandomState(0)
W = rng.rand(N, N) - 0.5
W *= 0.9 / np.max(np.abs(np.linalg.eigvals(W)))
res = Reservoir(W, rng.rand(N, ni) * 2 - 1, rng.rand(N, no) * 2 - 1, 0.005 * np.ones(ni), np.zeros(ni),
                5e-7 * np.ones(no), np.zeros(no), 0.001, True)
Wo = torch.randn(1, no, N + ni, device="cuda", dtype=dt) * 1e-6
us = torch.randn(B, T, ni, device="cuda", dtype=dt)
for _ in range(2):
    y = res.predict(us, Wo, transient=10, precision=prec, seed=3)
torch.cuda.synchronize()
print(float(y.abs().max()))
