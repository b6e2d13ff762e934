#!/usr/bin/env python
"""ms per predict call of the three tensor-core kernels at the small reservoirs (cfg2: 2x2, 100 neurons, 4096 frames; CDL
demo: 4x8, 300 neurons), T = 522."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "esn-ofdm-mimo_b200"))
from esn_b200 import Reservoir  # noqa: E402

T = 522
for N, ni, no, B in ((100, 4, 4, 4096), (100, 4, 4, 9472), (300, 16, 8, 9472), (512, 16, 8, 9472)):
    rng = np.random.RandomState(0)
    W = (rng.rand(N, N) - 0.5) * (0.9 * 2 / np.sqrt(N / 3))
    res = Reservoir(W, rng.rand(N, ni) * 2 - 1, rng.rand(N, no) * 2 - 1, 0.005 * np.ones(ni), np.zeros(ni),
                    5e-7 * np.ones(no), np.zeros(no), 0.001, True)
    x = torch.randn(B, T, ni, device="cuda")
    Wo = torch.randn(74, no, N + ni, dtype=torch.float64, device="cuda") * 1e-6
    gid = (torch.arange(B, device="cuda") // 128 % 74).to(torch.int32)
    out = []
    for path in ("tcr", "tc2", "tcs"):
        fn = lambda: res.predict(x, Wo, transient=10, group_ids=gid, precision=path, seed=1)  # noqa: E731
        fn()
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        out.append(f"{path} {best:6.2f} ms ({B / best / 1e3:5.2f} M sym/s)")
    print(f"N={N:4d} {ni}x{no} B={B:5d}: " + "  ".join(out))
