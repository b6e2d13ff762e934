#!/usr/bin/env python
"""Free-running fp64 prediction at cfg3 (512 neurons, T = 522, a readout per 128 frames): the streaming SIMT kernel
(ESN_HARVEST_DMMA=0) against the fp64 tensor-core kernel (default).  Run once per setting."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "esn-ofdm-mimo_b200"))
from esn_b200 import Reservoir  # noqa: E402

N, ni, no, T = 512, 16, 8, 522
B = int(sys.argv[1]) if len(sys.argv) > 1 else 9472
rng = np.random.RandomState(42)
W = (rng.rand(N, N) - 0.5) * (0.9 * 2 / np.sqrt(N / 3))
res = Reservoir(W, rng.rand(N, ni) * 2 - 1, rng.rand(N, no) * 2 - 1, 0.005 * np.ones(ni), np.zeros(ni),
                5e-7 * np.ones(no), np.zeros(no), 0.001, True)
u = torch.randn(B, T, ni, device="cuda", dtype=torch.float64)
G = (B + 127) // 128
W_out = torch.randn(G, no, N + ni, device="cuda", dtype=torch.float64) * 1e-6
gid = (torch.arange(B, device="cuda") // 128).to(torch.int32)
best = 1e9
for _ in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(4_000_000)
    a.record()
    y = res.predict(u, W_out, transient=10, group_ids=gid, precision="fp64", seed=1)
    b.record()
    torch.cuda.synchronize()
    best = min(best, a.elapsed_time(b))
flop = B * T * (2.0 * N * (N + ni + no) + 2.0 * no * (N + ni))
print(f"ESN_HARVEST_DMMA={os.environ.get('ESN_HARVEST_DMMA', 'auto')}: {B} frames {best:7.2f} ms  {B / best:7.1f} K symbols/s  {flop / best / 1e9:6.2f} TFLOP/s fp64")
