#!/usr/bin/env python
"""The SISO QPSK / AWGN demo loop (BASELINE.json configs[0]; reference
system_model_2/Demo_SISO_QPSK_AWGN_LDPC_ESN_with_ZF_LS.py:179-277, uncoded branch) driven through the
pyESN API exactly as the demo drives it -- one ESN per Eb/N0 built on numpy's GLOBAL generator, one
`fit` on the pilot symbol, one `predict` (continuation) per data symbol -- so the same file runs with the
reference's `pyESN` (CPU numpy) or with this repo's drop-in (B200):

  python examples/siso_qpsk_awgn.py                          # drop-in modules, GPU
  python examples/siso_qpsk_awgn.py --libs /path/to/reference/libs

The draw order of the global generator (channel, pilot, noise, reservoir weights, state noise, data bits)
is the demo's, so with equal seeds both runs see the same bits, channels and noise and their error counts
can be compared symbol for symbol (tests/test_gpu_chain.py::test_siso_demo_loop_matches_reference_counts).
"""
import argparse
import json
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def qpsk_table():
    """Unit-power 4-QAM in the demo's index order (idx = 2 i_re + i_im, levels -1, +1)."""
    lv = np.array([-1.0, 1.0])
    c = np.array([lv[i] + 1j * lv[q] for i in range(2) for q in range(2)])
    return c / math.sqrt(np.mean(np.abs(c) ** 2))


def run(esn_cls, ebno_db_list, n_symbols, n_res=200, N=512, seed=42, No=1e-5, clip_db=3.0, keep_first=False):
    """Returns {detector: [bit errors per Eb/N0]}, 'bits', per-symbol ESN errors, and (keep_first) the ESN's
    frequency-domain estimate of the first data symbol of every Eb/N0."""
    m = 2
    const = qpsk_table()
    pw = 2 ** np.arange(m)
    np.random.seed(seed)
    out = {k: [] for k in ("ESN", "MMSE", "ZF", "LS")}
    out.update(bits=[], esn_per_symbol=[], first_xhat=[], near_boundary=[])
    for ebno in ebno_db_list:
        Pi = 10 ** (ebno / 10) * No
        var_x = Pi * N
        A = math.sqrt(var_x) * 10 ** (clip_db / 20)
        amp = math.sqrt(Pi)

        def through_channel(X, h):
            x = N * np.fft.ifft(X) * amp                            # CP = 0
            x_pa = x / np.sqrt(1 + (np.abs(x) / A) ** 2)            # p_smooth = 1
            y = h * x_pa
            y = y + math.sqrt(len(y) * No / 2) * (np.random.randn(len(y)) + 1j * np.random.randn(len(y)))
            return x, y
        err = dict(ESN=0, MMSE=0, ZF=0, LS=0)
        per_symbol, near, bits = [], 0, 0
        for kk in range(1, n_symbols + 1):
            if kk == 1:                                             # pilot: flat unit-magnitude channel, LS estimate, ESN training
                h = np.random.randn() + 1j * np.random.randn()
                h /= abs(h)
                Xp = const[np.random.randint(0, 2 ** m, size=N)]
                xp, yp = through_channel(Xp, h)
                h_ls = np.mean((np.fft.fft(yp) / N) / (Xp * amp))
                esn = esn_cls(n_inputs=2, n_outputs=2, n_reservoir=n_res, spectral_radius=0.9, sparsity=0.1,
                              input_shift=np.zeros(2), input_scaling=(0.005 / math.sqrt(var_x)) * np.ones(2),
                              teacher_scaling=5e-7 * np.ones(2), teacher_shift=np.zeros(2),
                              feedback_scaling=np.zeros(2))
                esn.fit(np.column_stack([yp.real, yp.imag]), np.column_stack([xp.real, xp.imag]))
            tx = (np.random.rand(N * m) > 0.5).astype(np.int8)
            X = const[tx.reshape(N, m) @ pw]
            _, y = through_channel(X, h)
            Y = np.fft.fft(y) / N
            xh = esn.predict(np.column_stack([y.real, y.imag]))
            est = {"ESN": np.fft.fft(xh[:, 0] + 1j * xh[:, 1]) / N / amp,
                   "MMSE": np.conj(h) * Y / (abs(h) ** 2 + No / Pi + 1e-12) / amp,
                   "ZF": np.conj(h) * Y / (abs(h) ** 2 + 1e-12) / amp,
                   "LS": np.conj(h_ls) * Y / (abs(h_ls) ** 2 + 1e-12) / amp}
            for k, Xh in est.items():
                idx = np.argmin(np.abs(Xh[:, None] - const[None, :]), axis=1)
                rx = ((idx[:, None] >> np.arange(m)) & 1).astype(np.int8).reshape(-1)   # LSB first
                e = int(np.sum(rx != tx))
                err[k] += e
                if k == "ESN":
                    per_symbol.append(e)
                    near += int(np.sum((np.abs(Xh.real) < 1e-5) | (np.abs(Xh.imag) < 1e-5)))
                    if keep_first and kk == 1:
                        out["first_xhat"].append(Xh.copy())
            bits += N * m
        for k in err:
            out[k].append(err[k])
        out["bits"].append(bits)
        out["esn_per_symbol"].append(per_symbol)
        out["near_boundary"].append(near)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--libs", default=os.path.join(ROOT, "esn-ofdm-mimo_b200", "libs"),
                    help="directory holding pyESN.py (this repo's drop-in by default)")
    ap.add_argument("--ebno", default="0:3:30")
    ap.add_argument("--symbols", type=int, default=400)
    ap.add_argument("--nres", type=int, default=200)
    a = ap.parse_args()
    sys.path.insert(0, os.path.join(ROOT, "esn-ofdm-mimo_b200"))
    sys.path.insert(0, a.libs)
    from pyESN import ESN
    lo, st, hi = (float(x) for x in a.ebno.split(":"))
    ebno = list(np.arange(lo, hi + 1e-9, st))
    t0 = time.time()
    r = run(ESN, ebno, a.symbols, a.nres)
    dt = time.time() - t0
    print(json.dumps({"EbNo_dB": ebno, **{k: [e / b for e, b in zip(r[k], r["bits"])] for k in ("ESN", "MMSE", "ZF", "LS")},
                      "symbols_per_point": a.symbols, "seconds": round(dt, 2),
                      "ofdm_symbols_per_s": round(len(ebno) * a.symbols / dt, 1)}))


if __name__ == "__main__":
    main()
