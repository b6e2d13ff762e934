import os, sys
import numpy as np, torch
sys.path.insert(0, "/root/repo/esn-ofdm-mimo_b200")
from esn_b200 import Reservoir
from esn_b200._lib import ptr, check
import esn_b200.engine as E
N, ni, no, T = 512, 16, 8, 522
rng = np.random.RandomState(42)
W = rng.rand(N, N) - 0.5; W[rng.rand(N, N) < 0.1] = 0; W *= 0.9/np.max(np.abs(np.linalg.eigvals(W)))
res = Reservoir(W, rng.rand(N, ni)*2-1, rng.rand(N, no)*2-1, input_scaling=0.005*np.ones(ni), teacher_scaling=5e-7*np.ones(no), noise=0.001)
G = 1184
u = torch.randn(G, T, ni, device="cuda"); y = torch.randn(G, T, no, device="cuda")*1e-2
ext = res.harvest(u, y, precision="tc", seed=1)
def ev(): return torch.cuda.Event(enable_timing=True)
B, T_, P = ext.shape
m = T - 10
aff = res._aff[E.ESN_F64]
Gm = torch.empty((B, m, m), dtype=torch.float64, device="cuda")
rhs = torch.empty((B, m, no), dtype=torch.float64, device="cuda")
info = torch.zeros((B,), dtype=torch.int32, device="cuda")
Wout = torch.empty((B, no, P), dtype=torch.float64, device="cuda")
yy = y.contiguous()
for rep in range(2):
    e = [ev() for _ in range(4)]
    e[0].record()
    check(res.lib.esn_gram_f64(ptr(ext), E._CODE[ext.dtype], ptr(yy), E._CODE[yy.dtype], ptr(aff["t_scale"]), ptr(aff["t_shift"]), B, T, P, no, 10, 1, 0, 0, ptr(Gm), ptr(rhs), E._stream()), "gram")
    e[1].record()
    check(res.lib.esn_cholesky_solve_f64(ptr(Gm), ptr(rhs), B, m, no, ptr(info), E._stream()), "chol")
    e[2].record()
    check(res.lib.esn_readout_from_dual_f64(ptr(ext), E._CODE[ext.dtype], ptr(rhs), B, T, P, no, 10, ptr(Wout), E._stream()), "dual")
    e[3].record()
    torch.cuda.synchronize()
print("gram %.2f ms  cholesky+solve %.2f ms  dual readout %.2f ms  (G=%d)" % (e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), e[2].elapsed_time(e[3]), G))
