"""Readout-training throughput at cfg3 (4x8, N=512, T=522): harvest (fp64 / fp32 SIMT, tensor cores) and
Gram + Cholesky + readout, per batch of G pilots.  python profiles/fit_bench.py [G ...]"""
import os, sys, time
import numpy as np, torch
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "esn-ofdm-mimo_b200"))
from esn_b200 import Reservoir
N, ni, no, T = 512, 16, 8, 522
rng = np.random.RandomState(42)
W = rng.rand(N, N) - 0.5; W[rng.rand(N, N) < 0.1] = 0; W *= 0.9/np.max(np.abs(np.linalg.eigvals(W)))
res = Reservoir(W, rng.rand(N, ni)*2-1, rng.rand(N, no)*2-1, input_scaling=0.005*np.ones(ni), teacher_scaling=5e-7*np.ones(no), noise=0.001)
def ev(): return torch.cuda.Event(enable_timing=True)
for G in (int(a) for a in (sys.argv[1:] or (74, 592, 1184))):
    u = torch.randn(G, T, ni, device="cuda", dtype=torch.float64)
    y = torch.randn(G, T, no, device="cuda", dtype=torch.float64)*1e-2
    for prec in ("fp64", "fp32", "tc"):
        for rep in range(2):
            e=[ev() for _ in range(4)]
            e[0].record(); ext = res.harvest(u, y, precision=prec, seed=1); e[1].record()
            Wo, info = res.train_readout(ext, y, 10); e[2].record()
            torch.cuda.synchronize()
        print(f"G={G:5d} {prec}: harvest {e[0].elapsed_time(e[1]):8.2f} ms  gram+chol+readout {e[1].elapsed_time(e[2]):8.2f} ms  -> {G/ (e[0].elapsed_time(e[2])*1e-3):9.0f} fits/s  info_max {int(info.abs().max())}")
        del ext
