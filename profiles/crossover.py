#!/usr/bin/env python
"""Kernel time of one predict call (T = 522) for every explicit recurrence path on a batch x reservoir grid: the
table behind Reservoir.AUTO_TC_MIN_FRAMES / AUTO_TCS_MIN_FRAMES (esn_b200/engine.py: precision="auto").

    python profiles/crossover.py > profiles/r2_crossover.txt
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "esn-ofdm-mimo_b200")):
    sys.path.insert(0, p)
from esn_b200 import Reservoir  # noqa: E402
from oracle import esn_oracle as orc  # noqa: E402

T, TR = 522, 10


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def main():
    grid_n = [(100, 4, 4), (300, 4, 4), (512, 16, 8), (1024, 16, 8)]
    grid_b = [8, 32, 64, 128, 256, 512, 1024, 2048, 4736, 9472]
    print("ms per predict call (T = 522), best of 3; 'fp32' = esn_recurrence_run (cluster kernel at small batch, streaming SIMT above)")
    print("%6s %6s | %9s %9s %9s | %-6s %-6s" % ("N", "B", "fp32", "tc", "tcs", "best", "auto"))
    for n_res, n_in, n_out in grid_n:
        rng = np.random.RandomState(1)
        W, W_in, W_fb = orc.init_weights(rng, n_in, n_out, n_res, 0.9, 0.1)
        eng = Reservoir(W, W_in, W_fb, 0.005 * np.ones(n_in), np.zeros(n_in), 5e-7 * np.ones(n_out), np.zeros(n_out), 0.001, True)
        Wo = torch.from_numpy(rng.randn(1, n_out, n_res + n_in) * 1e-6).cuda()
        for B in grid_b:
            if n_res >= 1024 and B > 4736:
                continue
            u = torch.randn(B, T, n_in, device="cuda")
            t = {}
            if B <= 2048 or n_res <= 300:
                t["fp32"] = timed(lambda: eng.predict(u, Wo, transient=TR, precision="fp32", seed=1))
            if eng.tc_supported():
                rd = eng.tc_prepare(Wo, eng.input_scale_exponent(u))
                t["tc"] = timed(lambda: eng.predict_tc(u, rd, transient=TR, seed=1))
            rs = eng.tcs_prepare(Wo)
            t["tcs"] = timed(lambda: eng.predict_tcs(u, rs, transient=TR, seed=1))
            best = min(t, key=t.get)
            auto = eng.auto_predict_path(B, None)
            flag = "" if t.get(auto, 1e9) <= 1.1 * t[best] + 0.2 else "  <-- auto loses"
            print("%6d %6d | %9s %9s %9s | %-6s %-6s%s" % (
                n_res, B, *("%.2f" % t[k] if k in t else "-" for k in ("fp32", "tc", "tcs")), best, auto, flag))
            del u


if __name__ == "__main__":
    main()
