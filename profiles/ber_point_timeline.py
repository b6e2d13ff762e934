#!/usr/bin/env python
"""GPU timeline of one Monte-Carlo Eb/N0 point (linksim.ber_curve, 74 blocks x 128 frames): kernel start / end
per stream from the torch profiler's trace, gaps > 0.2 ms listed."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "esn-ofdm-mimo_b200"))
import esn_b200  # noqa: E402
from esn_b200 import Reservoir, linksim  # noqa: E402
from torch.profiler import profile, ProfilerActivity  # noqa: E402

esn_b200.load()
rng = np.random.RandomState(42)
N, ni, no = 512, 16, 8
W = rng.rand(N, N) - 0.5
W[rng.rand(N, N) < 0.1] = 0
W *= 0.9 / np.max(np.abs(np.linalg.eigvals(W)))
W_in, W_fb = rng.rand(N, ni) * 2 - 1, rng.rand(N, no) * 2 - 1


base = Reservoir(W, W_in, W_fb, 0.005 * np.ones(ni), np.zeros(ni), 5e-7 * np.ones(no), np.zeros(no), 0.001, True)


def factory(var_x):
    return base.rescaled(input_scaling=(0.005 / var_x ** 0.5) * np.ones(ni))


linksim.ber_curve(factory, 4, 8, 512, 4, [15], 74, 128, seed=1)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    linksim.ber_curve(factory, 4, 8, 512, 4, [15], 74, 128, seed=1)
    torch.cuda.synchronize()
path = "/tmp/ber_point_trace.json"
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e]
ev.sort(key=lambda e: e["ts"])
t0 = ev[0]["ts"]
print("span %.2f ms, %d GPU activities" % ((ev[-1]["ts"] + ev[-1]["dur"] - t0) / 1e3, len(ev)))
end = t0
for e in ev:
    gap = e["ts"] - end
    name = e["name"].replace("void (anonymous namespace)::", "").split("(")[0][:50]
    if e["dur"] > 150 or gap > 200 or (len(sys.argv) > 1 and float(sys.argv[1]) * 1e3 <= e["ts"] - t0 <= float(sys.argv[2]) * 1e3):
        print("%8.2f ms  +%6.2f  dur %7.3f ms  stream %-3s %s%s" % ((e["ts"] - t0) / 1e3, max(gap, 0) / 1e3, e["dur"] / 1e3,
              e.get("args", {}).get("stream", "?"), name, "   <-- idle gap before" if gap > 200 else ""))
    end = max(end, e["ts"] + e["dur"])
