import sys, os, time, math
import numpy as np, torch
sys.path.insert(0, "/root/repo/esn-ofdm-mimo_b200")
import esn_b200
from esn_b200 import Reservoir, linksim
esn_b200.load()
rng = np.random.RandomState(42)
N, ni, no = 512, 16, 8
W = rng.rand(N, N) - 0.5; W[rng.rand(N, N) < 0.1] = 0; W *= 0.9 / np.max(np.abs(np.linalg.eigvals(W)))
W_in, W_fb = rng.rand(N, ni) * 2 - 1, rng.rand(N, no) * 2 - 1
base = Reservoir(W, W_in, W_fb, 0.005 * np.ones(ni), np.zeros(ni), 5e-7 * np.ones(no), np.zeros(no), 0.001, True)


def factory(var_x):
    return base.rescaled(input_scaling=(0.005 / var_x ** 0.5) * np.ones(ni))
eb = [0, 15, 30]
linksim.ber_curve(factory, 4, 8, 512, 4, eb, 74, 128, seed=1)
torch.cuda.synchronize()
t0 = time.time()
c = linksim.ber_curve(factory, 4, 8, 512, 4, eb, 74, 128, seed=1)
torch.cuda.synchronize()
print("ber_curve: %.1f ms per Eb/N0 point of 9472 frames" % ((time.time() - t0) / len(eb) * 1e3), {k: c[k] for k in ("ESN", "MMSE")})
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    linksim.ber_curve(factory, 4, 8, 512, 4, [15], 74, 128, seed=1)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=60))
print(prof.key_averages().table(sort_by="cpu_time_total", row_limit=25, max_name_column_width=60))
