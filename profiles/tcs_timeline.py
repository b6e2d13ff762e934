#!/usr/bin/env python
"""Per-step phase timing of esn_predict_tcs from its own SM-clock stamps (esn_tcs_args.timeline): one CTA pair
alone, or with the whole GPU busy (--full: 9472 frames, stamps of CTA 0).

    python profiles/tcs_timeline.py [--nres 512] [--full] [--acc 2|4]
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "esn-ofdm-mimo_b200"))
from esn_b200 import Reservoir  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--nres", type=int, default=512)
ap.add_argument("--full", action="store_true")
ap.add_argument("--acc", type=int, default=0)
ap.add_argument("--ring-a", type=int, default=0)
ap.add_argument("--ring-b", type=int, default=0)
a = ap.parse_args()
N, ni, no, T = a.nres, 16, 8, 522
rng = np.random.RandomState(0)
W = (rng.rand(N, N) - 0.5) * (0.9 * 2 / np.sqrt(N / 3))
res = Reservoir(W, rng.rand(N, ni) * 2 - 1, rng.rand(N, no) * 2 - 1, 0.005 * np.ones(ni), np.zeros(ni),
                5e-7 * np.ones(no), np.zeros(no), 0.001, True)
B = 9472 if a.full else 128
G = 74
Wo = torch.randn(G, no, N + ni, device="cuda") * 1e-6
rd = res.tcs_prepare(Wo)
us = torch.randn(B, T, ni, device="cuda")
gid = ((torch.arange(B, device="cuda") // 18) % G).to(torch.int32)
tune = dict(accumulators=a.acc, ring_a=a.ring_a, ring_b=a.ring_b)
res.predict_tcs(us, rd, transient=10, group_ids=gid, seed=3, tune=tune)
tl = torch.zeros((T * 16 + 64 * 4,), dtype=torch.int64, device="cuda")
res.predict_tcs(us, rd, transient=10, group_ids=gid, seed=3, tune=tune, timeline=tl)
torch.cuda.synchronize()
tr = tl[T * 16:].cpu().numpy().reshape(64, 4)
t = tl[:T * 16].view(T, 16).cpu().numpy().astype(np.float64)[50:500]
t0 = t[:, 0:1]
step = np.diff(t[:, 0])
print(f"esn_predict_tcs N={N} B={B} tune={tune}: step {np.median(step):.0f} cycles (p10 {np.percentile(step, 10):.0f}, p90 {np.percentile(step, 90):.0f})")
rel = t - t0
cols = {0: "step start", 1: "pass0 TMEM free", 2: "pass0 state chunks issued", 3: "pass0 aug ready", 4: "pass0 committed",
        5: "pass1 TMEM free", 6: "pass1 state chunks issued", 8: "pass1 committed",
        9: "epi pass0 woke", 10: "epi pass0 TMEM drained", 11: "epi pass0 published", 12: "epi pass1 woke",
        13: "epi pass1 TMEM drained", 14: "epi pass1 published", 15: "frame warp: next aug published"}
for k in sorted(cols):
    v = rel[:, k]
    print(f"  +{np.median(v):9.0f}  {cols[k]}")
print("  per-chunk trace of step 200 (cycles from the first stamp): wait A | A landed | B landed | MMAs issued")
b0 = tr[0, 0]
for i in range(64):
    if tr[i, 3] == 0:
        break
    print("   item %2d: %7d %7d %7d %7d" % (i, tr[i, 0] - b0 if tr[i, 0] else -1, tr[i, 1] - b0 if tr[i, 1] else -1, tr[i, 2] - b0, tr[i, 3] - b0))
