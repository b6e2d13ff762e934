#!/usr/bin/env python
"""Kernel time of esn_recurrence_run for small batches: cluster kernel (weights resident in the cluster's
shared memory) vs the streaming kernel, predict at the cfg3 shape (N = 512, 16 in, 8 out, T = 522) and the
cfg2 shape (N = 100, 4 in, 4 out).  Picks the crossover behind esn_set_small_batch_limit's default."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "esn-ofdm-mimo_b200"))
from esn_b200 import Reservoir, _lib  # noqa: E402

lib = _lib.load()


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


for (N, ni, no) in ((512, 16, 8), (100, 4, 4)):
    rng = np.random.RandomState(0)
    W = rng.rand(N, N) - 0.5
    W *= 0.9 / np.max(np.abs(np.linalg.eigvals(W)))
    res = Reservoir(W, rng.rand(N, ni) * 2 - 1, rng.rand(N, no) * 2 - 1, 0.005 * np.ones(ni), np.zeros(ni),
                    5e-7 * np.ones(no), np.zeros(no), 0.001, True)
    T = 522
    for prec, dt in (("fp64", torch.float64), ("fp32", torch.float32)):
        Wo = torch.randn(1, no, N + ni, device="cuda", dtype=dt) * 1e-6
        print(f"N={N} {prec}:  B  cluster_ms  stream_ms  harvest_cluster_ms  harvest_stream_ms")
        for B in (1, 2, 4, 8, 16, 24, 32, 48, 64, 96, 128):
            us = torch.randn(B, T, ni, device="cuda", dtype=dt)
            ys = torch.randn(B, T, no, device="cuda", dtype=dt)
            row = []
            for mode in ("p", "h"):
                for lim in (1 << 20, 0):
                    lib.esn_set_small_batch_limit(lim)
                    if mode == "p":
                        row.append(timed(lambda: res.predict(us, Wo, transient=10, precision=prec, seed=3)))
                    else:
                        row.append(timed(lambda: res.harvest(us, ys, precision=prec, seed=3)))
            print(f"   {B:4d}  {row[0]:8.3f}  {row[1]:8.3f}  {row[2]:8.3f}  {row[3]:8.3f}")
