#!/usr/bin/env python
"""Calibration of the tensor-core paths' truncation-bias compensation (esn_tc_set_acc_k0).

The tensor core adds each K = 16 product block to its fp32 accumulator with truncation toward zero, so the
pre-activation W x + W_in u + W_fb y of an n-MMA chain comes out short by ~ n k0 relative, in the same direction
for every neuron; the recurrence amplifies that systematic shrink.  The epilogues multiply the accumulators by
1 + n k0.  This script sweeps k0 and prints, against the fp64 kernel on the same frames and the same noise:
  - the state error of a teacher-forced harvest (max |dx| / max |x| and rms),
  - the output error of a free-running predict with a readout trained on an fp64 harvest,
  - the fitted shrink of the states' pre-activations (least squares of atanh(x_tc) on atanh(x_64)) - 1,
for the resident kernel (tc) and the streamed-state kernel (tcs), at several reservoir sizes.

    python profiles/probes/tc_acc_bias.py > profiles/r2_tc_acc_bias.txt
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "esn-ofdm-mimo_b200"))
from esn_b200 import Reservoir  # noqa: E402
from esn_b200._lib import load  # noqa: E402

lib = load()
ni, no, T, tr = 16, 8, 522, 10
B = 128
K0S = [float(a) for a in os.environ.get("K0S", "0,1e-8,2e-8,3e-8,4e-8,5e-8,6e-8,8e-8").split(",")]
SIZES = [int(a) for a in os.environ.get("SIZES", "512,300,100,1024").split(",")]


def reservoir(N, seed=42):
    rng = np.random.RandomState(seed)
    W = rng.rand(N, N) - 0.5
    W[rng.rand(N, N) < 0.1] = 0
    if N <= 600:
        W *= 0.9 / np.max(np.abs(np.linalg.eigvals(W)))
    else:
        W *= 0.9 / (np.sqrt(N * 0.9 / 12.0))
    W_in, W_fb = rng.rand(N, ni) * 2 - 1, rng.rand(N, no) * 2 - 1
    return Reservoir(W, W_in, W_fb, 0.005 * np.ones(ni), np.zeros(ni), 5e-7 * np.ones(no), np.zeros(no), 0.001, True)


def io(seed):
    g = torch.Generator().manual_seed(seed)
    u = torch.randn(B, T, ni, generator=g, dtype=torch.float64)
    mix = torch.randn(3, ni, no, generator=g, dtype=torch.float64) / ni ** 0.5
    y = torch.zeros(B, T, no, dtype=torch.float64)
    for k in range(3):
        y[:, k:] += u[:, :T - k] @ mix[k]
    y += 0.01 * torch.randn(B, T, no, generator=g, dtype=torch.float64)
    return u.cuda(), y.cuda()


for N in SIZES:
    res = reservoir(N)
    u, y = io(N)
    uni_h = torch.rand(B, T - 1, N, dtype=torch.float32, generator=torch.Generator().manual_seed(7)).cuda()
    uni_p = torch.rand(B, T, N, dtype=torch.float32, generator=torch.Generator().manual_seed(8)).cuda()
    ext64 = res.harvest(u, y, precision="fp64", noise_uniforms=uni_h.double())
    x64 = ext64[:, :, :N]
    W_out, info = res.train_readout(ext64[:8], y[:8], tr)
    assert int(info.abs().max()) == 0
    gid = (torch.arange(B, device="cuda") // 64 % 8).to(torch.int32)
    y64 = res.predict(u, W_out, transient=tr, group_ids=gid, precision="fp64", noise_uniforms=uni_p.double())
    xs, ysc = float(x64.abs().max()), float(y64.abs().max())
    paths = (["tc", "tcr"] if N <= 512 else []) + ["tcs"]
    for path in paths:
        print(f"N={N} path={path}  (max|x| {xs:.3f}, max|y| {ysc:.3g})")
        for k0 in K0S:
            lib.esn_tc_set_acc_k0(k0)
            ext = res.harvest(u.float(), y.float(), precision=path, noise_uniforms=uni_h)
            x = ext[:, :, :N].double()
            dx = x - x64
            # shrink of the pre-activation: noise term removed first (x = tanh(z) + noise)
            nz = torch.zeros_like(x64)
            nz[:, 1:] = 0.001 * (uni_h.double() - 0.5)
            z64, ztc = torch.atanh((x64 - nz).clamp(-0.999999, 0.999999)), torch.atanh((x - nz).clamp(-0.999999, 0.999999))
            sel = z64.abs() > 0.05
            shrink = float((ztc[sel] * z64[sel]).sum() / (z64[sel] ** 2).sum()) - 1.0
            yp = res.predict(u.float(), W_out, transient=tr, group_ids=gid, precision=path, noise_uniforms=uni_p).double()
            dy = yp - y64
            print(f"  k0={k0:8.2e}: states max {float(dx.abs().max()) / xs:9.2e} rms {float(dx.pow(2).mean().sqrt()) / xs:9.2e}"
                  f"  z-shrink {shrink:+9.2e} | outputs max {float(dy.abs().max()) / ysc:9.2e} rms {float(dy.pow(2).mean().sqrt()) / ysc:9.2e}")
lib.esn_tc_set_acc_k0(-1.0)
