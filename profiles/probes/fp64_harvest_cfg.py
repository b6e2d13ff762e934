#!/usr/bin/env python
"""fp64 teacher-forced harvest (the parity-grade fit's first half) at cfg3 for the tile configurations of the streaming
SIMT kernel (ESN_SIMT_CFG=BT,WN,KC is read once per process: this script re-executes itself per configuration).
    python profiles/probes/fp64_harvest_cfg.py [pilots]"""
import os
import subprocess
import sys

if "ESN_SIMT_CFG_RUN" not in os.environ:
    B = sys.argv[1] if len(sys.argv) > 1 else "4736"
    for cfg in ("", "32,2,8", "16,4,8", "8,4,8", "8,2,8", "8,4,4"):
        env = dict(os.environ, ESN_SIMT_CFG_RUN="1")
        if cfg:
            env["ESN_SIMT_CFG"] = cfg
        r = subprocess.run([sys.executable, __file__, B], env=env, capture_output=True, text=True)
        print(f"cfg {cfg or 'auto':8s}: {r.stdout.strip() or r.stderr.strip()[-200:]}")
    sys.exit(0)

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "esn-ofdm-mimo_b200"))
from esn_b200 import Reservoir  # noqa: E402

B = int(sys.argv[1])
N, ni, no, T = 512, 16, 8, 522
rng = np.random.RandomState(42)
W = (rng.rand(N, N) - 0.5) * (0.9 * 2 / np.sqrt(N / 3))
res = Reservoir(W, rng.rand(N, ni) * 2 - 1, rng.rand(N, no) * 2 - 1, 0.005 * np.ones(ni), np.zeros(ni),
                5e-7 * np.ones(no), np.zeros(no), 0.001, True)
res.lib.esn_set_small_batch_limit(0)
u = torch.randn(B, T, ni, device="cuda", dtype=torch.float64)
y = torch.randn(B, T, no, device="cuda", dtype=torch.float64)
res.harvest(u, y, precision="fp64", seed=1)
torch.cuda.synchronize()
best = 1e9
for _ in range(2):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    res.harvest(u, y, precision="fp64", seed=1)
    b.record()
    torch.cuda.synchronize()
    best = min(best, a.elapsed_time(b))
flop = B * (T - 1) * 2 * N * (N + ni + no)
print(f"{best:8.2f} ms  {flop / best / 1e9:6.2f} TFLOP/s fp64  ({B} pilots)")
