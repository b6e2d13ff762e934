#!/usr/bin/env python
"""Predict throughput for the sweep's large reservoirs (cfg4: 600 / 1024 / 2048 neurons, 4x8, T = 522), device-
resident inputs: the streamed-state tensor-core kernel (esn_predict_tcs) beside the streaming SIMT kernel (fp32).
TFLOP/s = algorithmic (2 N (N + n_in + n_out) + 2 n_out (N + n_in) per step and frame); the in-kernel
issuer stamps give cycles per time step of one CTA pair (128 frames).

    python profiles/large_n_bench.py [--simt] > profiles/r2_large_n_tcs.txt
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "esn-ofdm-mimo_b200"))
from esn_b200 import Reservoir  # noqa: E402

ni, no, T = 16, 8, 522
simt = "--simt" in sys.argv
for N in (512, 600, 1024, 2048):
    rng = np.random.RandomState(0)
    W = (rng.rand(N, N) - 0.5) * (0.9 * 2 / np.sqrt(N / 3))          # ~ spectral radius 0.9 without the eigen solve
    res = Reservoir(W, rng.rand(N, ni) * 2 - 1, rng.rand(N, no) * 2 - 1, 0.005 * np.ones(ni), np.zeros(ni),
                    5e-7 * np.ones(no), np.zeros(no), 0.001, True)
    G = 74
    Wo = torch.randn(G, no, N + ni, device="cuda") * 1e-6
    rd = res.tcs_prepare(Wo)
    for B in (2368, 9472):
        us = torch.randn(B, T, ni, device="cuda")
        gid = ((torch.arange(B, device="cuda") // 18) % G).to(torch.int32)       # the reference's L = 19 cadence
        paths = [("tcs", lambda: res.predict_tcs(us, rd, transient=10, group_ids=gid, seed=3))]
        if simt and B == 2368:
            paths.append(("simt", lambda: res.predict(us, Wo.float(), transient=10, group_ids=gid, precision="fp32", seed=3)))
        for name, fn in paths:
            fn()
            torch.cuda.synchronize()
            best = 1e9
            for _ in range(2):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            flop = B * T * (2 * N * (N + ni + no) + 2 * no * (N + ni))
            extra = ""
            if name == "tcs":
                tl = torch.zeros((T * 16 + 256,), dtype=torch.int64, device="cuda")
                res.predict_tcs(us[:128], rd, transient=10, group_ids=gid[:128], seed=3, timeline=tl)
                torch.cuda.synchronize()
                st = tl[:T * 16].view(T, 16)[:, 0].cpu().numpy()
                extra = "  %.1f K cycles/step alone" % (np.median(np.diff(st[50:])) / 1e3)
            print(f"N={N:5d} B={B:5d} {name:5s}: {best:9.2f} ms  {B / best * 1e3:10.0f} sym/s  {flop / best / 1e9:7.2f} TFLOP/s{extra}")
