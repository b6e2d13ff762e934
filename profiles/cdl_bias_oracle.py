#!/usr/bin/env python
"""CPU-oracle run of the CDL demo's configuration (4x8, 128 subcarriers, 300 neurons, TDL-B taps, 16-QAM,
L = 75: 1 pilot + 74 data symbols per coherence block; system_model_2/Demo_MIMO_4x8_Sionna_CDL_ESN_v2.py) to
explain the gap between this repo's device curve and the BERs the reference published for that script
(results/results_4x8_cdl_coded_uncoded/CDLB_run_01/results_ber.csv: 1000 OFDM symbols = 13 channel draws per
point).  Two variants of the float64 oracle on identical channels, bits and noise:
  fixedW : ONE reservoir (random_state 42) for all blocks, a readout per block -- what the device loop does
  freshW : a new random reservoir per block -- what the reference's demo does (unseeded ESN per pilot)
Prints mean BER +- standard error over blocks, and the spread of 13-block sub-samples.

    python profiles/cdl_bias_oracle.py [--blocks 208] > profiles/r2_cdl_bias_oracle.txt
"""
import argparse
import math
import multiprocessing as mp
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import esn_oracle as orc  # noqa: E402

N, N_t, N_r, m, isi, n_res, per = 128, 4, 8, 4, 8, 300, 74
PUBLISHED = {0: (0.3904, 0.3196), 6: (0.3231, 0.1854), 12: (0.2445, 0.0786), 18: (0.1860, 0.0345), 24: (0.1591, 0.0219),
             30: (0.1569, 0.0189)}
_FIXED = {}


def one_block(args):
    ebno, seed = args
    from threadpoolctl import threadpool_limits
    with threadpool_limits(limits=1):
        blk = orc.synth_block(seed, N, N_t, N_r, m, ebno, per, isi_duration=isi, channel="tdlb")
        cp, Pi, No = isi - 1, blk["Pi"], blk["No"]
        maxd = int(math.ceil(isi / 2) + 2)
        kw = dict(n_inputs=2 * N_r, n_outputs=2 * N_t, n_reservoir=n_res, spectral_radius=0.9, sparsity=0.1,
                  input_shift=np.zeros(2 * N_r), input_scaling=(0.005 / blk["var_x"] ** 0.5) * np.ones(2 * N_r),
                  teacher_scaling=5e-7 * np.ones(2 * N_t), teacher_shift=np.zeros(2 * N_t))
        if "W" not in _FIXED:
            e0 = orc.OracleESN(random_state=42, **kw)
            _FIXED.update(W=e0.W, W_in=e0.W_in, W_feedb=e0.W_feedb)
        out = {}
        for variant in ("fixedW", "freshW"):
            esn = orc.OracleESN(random_state=np.random.RandomState(100000 + seed), **kw)      # freshW: its own weights
            if variant == "fixedW":
                esn.W, esn.W_in, esn.W_feedb = _FIXED["W"], _FIXED["W_in"], _FIXED["W_feedb"]
            r = orc.train_generic(esn, 0, 0, maxd, cp, N, N_t, N_r, isi, blk["pilot"]["y_CP"], blk["pilot"]["x_CP"])
            d, nforget = int(r[6]), int(r[7])
            errs = 0
            for f in blk["data"]:
                X = orc.esn_output_to_freq(esn.predict(orc.pack_rx(f["y_CP"], d), nforget, continuation=False), N, N_t, Pi)
                errs += int((orc.indices_to_bits(orc.hard_demap_indices(X, blk["const"]), m) != f["bits"]).sum())
            out[variant] = errs
        Y_LS = orc.rx_fft(blk["pilot"]["y_LS_CP"], cp, N)
        H_LS, H_MM = orc.channel_estimate(Y_LS, blk["pilot"]["X_LS"], Pi, No, N, N_t, N_r, blk["isi_magnitude"], isi)
        errs = 0
        for f in blk["data"]:
            X = orc.equalize(orc.rx_fft(f["y_CP"], cp, N), H_MM, math.sqrt(Pi), No / Pi)
            errs += int((orc.indices_to_bits(orc.hard_demap_indices(X, blk["const"]), m) != f["bits"]).sum())
        out["MMSE"] = errs
        return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--blocks", type=int, default=208)
    ap.add_argument("--ebno", type=int, nargs="+", default=[12, 18, 24, 30])
    a = ap.parse_args()
    bits = per * N * N_t * m
    print(f"CDL demo configuration on the CPU oracle: {a.blocks} coherence blocks x {per} data symbols per Eb/N0")
    print("%5s | %-22s %-22s %-22s | published ESN / MMSE (13 blocks) | sd of a 13-block mean (fixedW)" %
          ("EbN0", "ESN fixedW", "ESN freshW", "MMSE"))
    with mp.get_context("fork").Pool(os.cpu_count() or 1) as pool:
        for ebno in a.ebno:
            res = pool.map(one_block, [(ebno, 7000 + 13 * ebno + b) for b in range(a.blocks)], chunksize=2)
            col = {}
            for k in ("fixedW", "freshW", "MMSE"):
                v = np.array([r[k] for r in res]) / bits
                col[k] = (v.mean(), v.std(ddof=1) / math.sqrt(len(v)), v.std(ddof=1))
            sd13 = col["fixedW"][2] / math.sqrt(13)
            print("%5d | %.4f +- %.4f       %.4f +- %.4f       %.4f +- %.4f       | %.4f / %.4f                  | %.4f" %
                  (ebno, col["fixedW"][0], col["fixedW"][1], col["freshW"][0], col["freshW"][1], col["MMSE"][0], col["MMSE"][1],
                   PUBLISHED[ebno][0], PUBLISHED[ebno][1], sd13))
            sys.stdout.flush()


if __name__ == "__main__":
    main()
