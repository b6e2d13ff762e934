#!/usr/bin/env python
"""fp64 teacher-forced harvest of a GPU-filling pilot batch at cfg3 (4736 pilots, 512 neurons, T = 522): the streaming
SIMT kernel (ESN_HARVEST_DMMA=0) against the fp64 tensor-core kernel (default for such batches).  Run once per setting:
the switch is read when the library first dispatches a harvest."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "esn-ofdm-mimo_b200"))
from esn_b200 import Reservoir  # noqa: E402

N, ni, no, T = 512, 16, 8, 522
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4736
rng = np.random.RandomState(42)
W = (rng.rand(N, N) - 0.5) * (0.9 * 2 / np.sqrt(N / 3))
res = Reservoir(W, rng.rand(N, ni) * 2 - 1, rng.rand(N, no) * 2 - 1, 0.005 * np.ones(ni), np.zeros(ni),
                5e-7 * np.ones(no), np.zeros(no), 0.001, True)
u = torch.randn(B, T, ni, device="cuda", dtype=torch.float64)
y = torch.randn(B, T, no, device="cuda", dtype=torch.float64) * 1e-2
best = 1e9
for _ in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(4_000_000)
    a.record()
    ext = res.harvest(u, y, precision="fp64", seed=1)
    b.record()
    torch.cuda.synchronize()
    best = min(best, a.elapsed_time(b))
    del ext
flop = B * (T - 1) * 2.0 * N * (N + ni + no)
print(f"ESN_HARVEST_DMMA={os.environ.get('ESN_HARVEST_DMMA', 'auto')}: {B} pilots {best:7.2f} ms  {flop / best / 1e9:6.2f} TFLOP/s fp64")
