#!/usr/bin/env python
"""Pinned host -> device copy rate of this box with the GPU idle and while the recurrence kernel runs on another stream
(no dependency between the two): is the end-to-end step bound by the box's DMA rate under load?"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "esn-ofdm-mimo_b200"))
from esn_b200 import Reservoir  # noqa: E402

N, ni, no, T, B = 512, 16, 8, 522, 9472
rng = np.random.RandomState(42)
W = (rng.rand(N, N) - 0.5) * (0.9 * 2 / np.sqrt(N / 3))
res = Reservoir(W, rng.rand(N, ni) * 2 - 1, rng.rand(N, no) * 2 - 1, 0.005 * np.ones(ni), np.zeros(ni),
                5e-7 * np.ones(no), np.zeros(no), 0.001, True)
x = torch.randn(B, T, ni, device="cuda")
rd = res.tcs_prepare(torch.randn(1, no, N + ni, dtype=torch.float64, device="cuda") * 1e-6)
su = res.input_scale_exponent(x)
h = torch.empty((B, T, ni), dtype=torch.float32).pin_memory()
d = [torch.empty_like(x), torch.empty_like(x)]
cs = [torch.cuda.Stream(), torch.cuda.Stream()]
ks = torch.cuda.Stream()
half = B // 2


def copies(n, split):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record(cs[0])
    for k in range(n):
        if split:
            for s, part in zip(cs, (slice(0, half), slice(half, B))):
                with torch.cuda.stream(s):
                    d[k & 1][part].copy_(h[part], non_blocking=True)
        else:
            with torch.cuda.stream(cs[0]):
                d[k & 1].copy_(h, non_blocking=True)
    cs[0].wait_stream(cs[1])
    b.record(cs[0])
    torch.cuda.synchronize()
    return h.numel() * 4 * n / (a.elapsed_time(b) * 1e-3) / 1e9


for split in (False, True):
    idle = copies(8, split)
    with torch.cuda.stream(ks):
        for _ in range(24):
            res.predict_tcr(x, rd, transient=10, seed=1, su_exp=su)
    busy = copies(8, split)
    torch.cuda.synchronize()
    print(f"{'two streams' if split else 'one stream '}: idle {idle:5.1f} GB/s, while esn_recur_tcr runs {busy:5.1f} GB/s")

# and the other way round: the kernel's time with and without copies in flight
def kernels(n, with_copies):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    if with_copies:
        for k in range(3 * n):
            with torch.cuda.stream(cs[k & 1]):
                d[k & 1].copy_(h, non_blocking=True)
    with torch.cuda.stream(ks):
        a.record(ks)
        for _ in range(n):
            res.predict_tcr(x, rd, transient=10, seed=1, su_exp=su)
        b.record(ks)
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


for _ in range(2):
    print(f"esn_recur_tcr: {kernels(12, False):.2f} ms per launch alone, {kernels(12, True):.2f} ms with H2D copies in flight")
