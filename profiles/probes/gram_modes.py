#!/usr/bin/env python
"""Gram kernel (syrk_dmma) by mode at cfg3, 1184 pilots of fp32 extended states: dual per pilot (512 x 512), primal per
pilot (528 x 528), primal summed over the pilots (the shared readout)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "esn-ofdm-mimo_b200"))
from esn_b200 import Reservoir  # noqa: E402
from esn_b200._lib import check, ptr  # noqa: E402
import esn_b200.engine as E  # noqa: E402

N, ni, no, T, G, tr = 512, 16, 8, 522, 1184, 10
rng = np.random.RandomState(42)
W = (rng.rand(N, N) - 0.5) * (0.9 * 2 / np.sqrt(N / 3))
res = Reservoir(W, rng.rand(N, ni) * 2 - 1, rng.rand(N, no) * 2 - 1, 0.005 * np.ones(ni), np.zeros(ni),
                5e-7 * np.ones(no), np.zeros(no), 0.001, True)
u = torch.randn(G, T, ni, device="cuda")
y = torch.randn(G, T, no, device="cuda") * 1e-2
ext = res.harvest(u, y, precision="tc", seed=1)
P, m = N + ni, T - tr
aff = res._aff[E.ESN_F64]


def run(dual, shared):
    n = m if dual else P
    nprob = 1 if shared else G
    Gm = torch.empty((nprob, n, n), dtype=torch.float64, device="cuda")
    rhs = torch.empty((nprob, n if dual else P, no), dtype=torch.float64, device="cuda")
    best = 1e9
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        check(res.lib.esn_gram_f64(ptr(ext), E._CODE[ext.dtype], ptr(y), E._CODE[y.dtype], ptr(aff["t_scale"]), ptr(aff["t_shift"]),
                                   G, T, P, no, tr, int(dual), int(shared), 0, ptr(Gm), ptr(rhs), E._stream()), "gram")
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    flop = G * n * (n + 1) * (P if dual else m)              # lower triangle, 2 flop per MAC
    print(f"dual={dual} shared={shared}: {best:7.2f} ms  {flop / best / 1e9:6.2f} TFLOP/s fp64 (lower triangle)")


run(1, 0)
run(0, 0)
run(0, 1)
