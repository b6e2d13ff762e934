#!/usr/bin/env python
"""One predict launch per tensor-core kernel at cfg3 (for ncu captures):  python profiles/probes/tc_once.py tcr|tc|tcs [B]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "esn-ofdm-mimo_b200"))
from esn_b200 import Reservoir  # noqa: E402

path = sys.argv[1] if len(sys.argv) > 1 else "tcr"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 9472
N, ni, no, T = 512, 16, 8, 522
rng = np.random.RandomState(42)
W = (rng.rand(N, N) - 0.5) * (0.9 * 2 / np.sqrt(N / 3))
res = Reservoir(W, rng.rand(N, ni) * 2 - 1, rng.rand(N, no) * 2 - 1, 0.005 * np.ones(ni), np.zeros(ni),
                5e-7 * np.ones(no), np.zeros(no), 0.001, True)
x = torch.randn(B, T, ni, device="cuda")
Wo = torch.randn(74, no, N + ni, dtype=torch.float64, device="cuda") * 1e-6
gid = (torch.arange(B, device="cuda") // 128 % 74).to(torch.int32)
for _ in range(2):
    y = res.predict(x, Wo, transient=10, group_ids=gid, precision=path, seed=1)
torch.cuda.synchronize()
print(path, float(y.abs().max()))
