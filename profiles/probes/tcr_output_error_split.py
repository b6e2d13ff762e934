#!/usr/bin/env python
"""Where does the output error of esn_recur_tcr on link frames come from -- the states or the fp32 readout?
Free-running predict on frames of the block-fading 4x8 link with pilot-trained readouts (bench.py's generator, two
blocks): outputs of the kernel vs the fp64 kernel, and the fp64 readout applied to the kernel's own states."""
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "esn-ofdm-mimo_b200"))
import esn_b200  # noqa: E402
from esn_b200 import Reservoir, linksim  # noqa: E402

dev = "cuda"
G, per, N, N_t, N_r, cp, m, isi, d = 2, 128, 512, 4, 8, 7, 4, 8, 3
B, T = G * per, N + cp + d
No, ebno = 1e-5, 15.0
Pi = 10 ** (ebno / 10) * No
A = math.sqrt(Pi * N) * 10 ** (3 / 20)
std = math.sqrt((N + cp) * No / 2)
gen = torch.Generator(device=dev).manual_seed(1)
rng = np.random.RandomState(42)
W = rng.rand(N, N) - 0.5
W[rng.rand(N, N) < 0.1] = 0
W *= 0.9 / np.max(np.abs(np.linalg.eigvals(W)))
res = Reservoir(W, rng.rand(N, 2 * N_r) * 2 - 1, rng.rand(N, 2 * N_t) * 2 - 1, (0.005 / math.sqrt(Pi * N)) * np.ones(2 * N_r),
                np.zeros(2 * N_r), 5e-7 * np.ones(2 * N_t), np.zeros(2 * N_t), 0.001, True)
taps = (torch.randn((G, N_r, N_t, isi), generator=gen, device=dev, dtype=torch.float64)
        + 1j * torch.randn((G, N_r, N_t, isi), generator=gen, device=dev, dtype=torch.float64)) / math.sqrt(2)
taps = taps * linksim.isi_profile(isi, dev).sqrt()
pil_idx = torch.randint(0, 16, (G, N, N_t), generator=gen, device=dev, dtype=torch.uint8)
pil = esn_b200.ofdm.synth_frames(pil_idx, taps, Pi, A, N, cp, m, std, delay=d, seed=11, dtype=torch.float64, want_x_cp=True, want_y_cp=False)
pil_y = torch.zeros((G, T, 2 * N_t), dtype=torch.float64, device=dev)
pil_y[:, d:, :] = torch.view_as_real(pil["x_cp"]).reshape(G, N + cp, 2 * N_t)
ext = res.harvest(pil["esn_in"], pil_y, precision="fp64", seed=7)
W_out, info = res.train_readout(ext, pil_y, d + cp)
gid = (torch.arange(B, device=dev) // per).to(torch.int32)
tx = torch.randint(0, 16, (B, N, N_t), generator=gen, device=dev, dtype=torch.uint8)
fr = esn_b200.ofdm.synth_frames(tx, taps.to(torch.complex64), Pi, A, N, cp, m, std, delay=d, chan_index=gid, seed=23,
                                dtype=torch.float32, want_y_cp=False)["esn_in"]
y64, e64 = res.predict(fr.double(), W_out, transient=d + cp, group_ids=gid, precision="fp64", seed=99, return_ext=True)
for path in ("tcr", "tc2", "tcs"):
    y, e = res.predict(fr, W_out, transient=d + cp, group_ids=gid, precision=path, seed=99, return_ext=True)
    # fp64 readout on the kernel's own states (teacher units: (W_out [x; u]) / t_scale)
    y_st = torch.einsum("btp,bop->bto", e.double(), W_out[gid.long()])[:, d + cp:, :] / 5e-7
    n = float(y64.norm())
    print(f"{path}: states {float((e.double() - e64).norm() / e64.norm()):.2e} | outputs {float((y.double() - y64).norm()) / n:.2e} | "
          f"fp64 readout of its states {float((y_st - y64).norm()) / n:.2e} | kernel readout vs fp64 readout of the same states "
          f"{float((y.double() - y_st).norm()) / n:.2e}")
