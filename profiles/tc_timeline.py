#!/usr/bin/env python
"""Per-step phase timing of the tensor-core recurrence kernel (CTA 0), from the
kernel's own SM-clock stamps (esn_tc_predict_args.timeline).  Run on the GPU box:
    python profiles/tc_timeline.py [frames]
Rows (cycles): the issuer's waits for the state barriers (tA0: chunks 0, 2 rewritten and every G0
accumulator read; tB: everything rewritten), its issue phases (ring-bound when the data is late) and
the epilogue of warp 4 (first 16-neuron block, then the rest)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "esn-ofdm-mimo_b200"))
from esn_b200 import Reservoir  # noqa: E402

TCR = "--tcr" in sys.argv          # the resident kernel with the CUDA-core readout (esn_recur_tcr) instead of esn_predict_tc2
argv = [a for a in sys.argv[1:] if not a.startswith("--")]
B = int(argv[0]) if argv else 148 * 64
N, ni, no, T = int(os.environ.get("TL_N", "512")), 16, 8, 522
rng = np.random.RandomState(42)
W = rng.rand(N, N) - 0.5
W[rng.rand(N, N) < 0.1] = 0
W *= 0.9 / np.max(np.abs(np.linalg.eigvals(W)))
res = Reservoir(W, rng.rand(N, ni) * 2 - 1, rng.rand(N, no) * 2 - 1, input_scaling=0.005 * np.ones(ni),
                teacher_scaling=5e-7 * np.ones(no), noise=0.001)
x = torch.randn(B, T, ni, device="cuda")
Wo = torch.randn(1, no, N + ni, dtype=torch.float64, device="cuda") * 1e-6
rd = res.tcs_prepare(Wo) if TCR else res.tc_prepare(Wo, res.input_scale_exponent(x))
tl = torch.zeros(T + 1 + 32, 8, dtype=torch.int64, device="cuda")   # + per-item trace of step 200
for _ in range(2):
    if TCR:
        res.predict_tcr(x, rd, transient=10, timeline=tl)
    else:
        res.predict_tc(x, rd, transient=10, timeline=tl)
torch.cuda.synchronize()
full = tl.cpu().numpy()
t = full[50:500]
step = t[1:, 0] - t[:-1, 0]
if TCR:
    # esn_recur_tcr: G0 finishes before G1 (commit d0 / d1); stamp 7 = G0 chunks 4-7 issued, stamp 4 = epilogue woke on d0
    rows = (("step", step),
            ("wait tA0 (chunks 0,2 rewritten)", t[:, 1] - t[:, 0]),
            ("issue G0 c0-3, G1 c0-3", t[:, 2] - t[:, 1]),
            ("wait tB (all state rewritten)", t[:, 3] - t[:, 2]),
            ("issue G0 c4-7", t[:, 7] - t[:, 3]),
            ("wait y, aug G0, G1 c4-7 + aug (to next step)", t[1:, 0] - t[:-1, 7]),
            ("epilogue G0 block a (+publish)", t[:, 5] - t[:, 4]),
            ("epilogue G0 b, wait d1, G1, readout sweep", t[:, 6] - t[:, 5]))
else:
  rows = (("step", step),
          ("wait tA0 (chunks 0,2 rewritten)", t[:, 1] - t[:, 0]),
          ("issue G0 c0-3, G1 c0-3", t[:, 2] - t[:, 1]),
          ("wait tB (all state rewritten)", t[:, 3] - t[:, 2]),
          ("issue G1 c4-7 (y), G0 c4-6", t[:, 7] - t[:, 3]),
          ("wait y, aug, G0 c7 -> D ready", t[:, 4] - t[:, 7]),
          ("epilogue block a (+publish)", t[:, 5] - t[:, 4]),
          ("epilogue block b + G1", t[:, 6] - t[:, 5]))
for name, v in rows:
    print(f"{name:32s} mean {v.mean():9.0f}  p10 {np.percentile(v, 10):9.0f}  p90 {np.percentile(v, 90):9.0f} cycles")

tr = full[T + 1:].reshape(-1, 4)[:56]
if tr[:, 2].any():
    t0 = tr[0, 0] if tr[0, 0] else tr[0, 2]
    print("per-item trace of step 200 (cycles from the first stamp): slot_free_seen  copy_issued | data_seen  mmas_issued")
    for i, r in enumerate(tr):
        if r[2] == 0:
            break
        print(f"  item {i:2d}: {r[0]-t0:8d} {r[1]-t0:8d} | {r[2]-t0:8d} {r[3]-t0:8d}")
