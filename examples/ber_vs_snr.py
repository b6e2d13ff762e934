#!/usr/bin/env python
"""BER-vs-SNR Monte-Carlo of the block-fading template on the GPU(s), written in the reference's result
formats (results_ber.csv / results_ber.pkl; see esn_b200/results.py).

  python examples/ber_vs_snr.py --out results_4x8/run_01                      # one GPU
  torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 examples/ber_vs_snr.py --out ...

Per Eb/N0 point: `--blocks` coherence blocks (sharded over ranks), each with one pilot symbol (readout
training + LS / MMSE channel estimates) and `--frames-per-block` data symbols detected by the ESN and by
the Perfect-ZF / LS-ZF / MMSE baselines; error counters are summed over ranks with one allreduce.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "esn-ofdm-mimo_b200"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nt", type=int, default=4)
    ap.add_argument("--nr", type=int, default=8)
    ap.add_argument("--nsub", type=int, default=512)
    ap.add_argument("--nres", type=int, default=512)
    ap.add_argument("--qam-bits", type=int, default=4)
    ap.add_argument("--ebno", default="0:3:30", help="start:step:stop (inclusive) in dB")
    ap.add_argument("--blocks", type=int, default=74)
    ap.add_argument("--frames-per-block", type=int, default=128)
    ap.add_argument("--fit-precision", default="fp64", choices=["fp64", "fp32", "tc"])
    ap.add_argument("--detect-precision", default="tc", choices=["tc", "auto", "fp32", "fp64"])
    ap.add_argument("--train-fixed-ebno", type=float, default=None,
                    help="also train the template's second ESN on the pilot re-sent at this Eb/N0 (12 dB in the demos)")
    ap.add_argument("--channel", default="rayleigh", choices=["rayleigh", "tdlb"],
                    help="rayleigh: exponential 8-tap profile of the NBF template; tdlb: TDL-B taps of the CDL demo")
    ap.add_argument("--blocks-per-launch", type=int, default=0,
                    help="coherence blocks processed together (0 = about 75 K frames per launch; memory bound: ~17 MB per block of 128 frames at 4x8, N = 512); "
                         "with --fit-precision tc the pilot harvest costs the same 10 ms for up to 9472 blocks")
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--out", default="results_ber_run")
    a = ap.parse_args()

    import torch
    import esn_b200
    from esn_b200 import Reservoir, dist as D, linksim, results
    rank, world, local = D.init_from_env()
    torch.cuda.set_device(local)
    esn_b200.load()
    lo, st, hi = (float(x) for x in a.ebno.split(":"))
    ebno = list(np.arange(lo, hi + 1e-9, st))
    rng = np.random.RandomState(a.seed)                   # pyESN.initweights order (libs/pyESN.py:93-109)
    N, ni, no = a.nres, 2 * a.nr, 2 * a.nt
    W = rng.rand(N, N) - 0.5
    W[rng.rand(N, N) < 0.1] = 0
    W *= 0.9 / np.max(np.abs(np.linalg.eigvals(W)))
    W_in, W_fb = rng.rand(N, ni) * 2 - 1, rng.rand(N, no) * 2 - 1

    base = Reservoir(W, W_in, W_fb, 0.005 * np.ones(ni), np.zeros(ni), 5e-7 * np.ones(no), np.zeros(no), 0.001, True)

    def factory(var_x):                                   # same device weights, inputs rescaled per Eb/N0 (:237-241)
        return base.rescaled(input_scaling=(0.005 / var_x ** 0.5) * np.ones(ni))
    warm = torch.zeros(1, device="cuda")
    D.allreduce_sum_(warm)                                # NCCL communicator set-up stays out of the timing
    torch.cuda.synchronize()
    t0 = time.time()
    c = linksim.ber_curve(factory, a.nt, a.nr, a.nsub, a.qam_bits, ebno, a.blocks, a.frames_per_block,
                          seed=a.seed, fit_precision=a.fit_precision, detect_precision=a.detect_precision,
                          channel=a.channel, max_blocks_per_launch=a.blocks_per_launch,
                          train_fixed_ebno_db=a.train_fixed_ebno)
    torch.cuda.synchronize()
    dt = time.time() - t0
    if rank == 0:
        os.makedirs(a.out, exist_ok=True)
        results.write_results_csv(os.path.join(a.out, "results_ber.csv"), c["EBN0"], c["ESN"], c["MMSE"])
        meta = {"N": a.nsub, "N_t": a.nt, "N_r": a.nr, "IsiDuration": 8, "CP": 7,
                "NumOfdmSymbols": a.blocks * (a.frames_per_block + 1),
                "esn": {"n_reservoir": a.nres, "spectral_radius": 0.9, "input_scaler": 0.005, "teacher_scaling_base": 5e-7},
                "channel": {"model": "CDL-B (TDL-equivalent)" if a.channel == "tdlb" else "block-fading Rayleigh, 8 taps, exponential profile"},
                "all_detectors": {k: c[k] for k in c if not k.startswith("_") and k != "EBN0"}, "gpus": world, "seconds": dt}
        results.write_results_pkl(os.path.join(a.out, "results_ber.pkl"),
                                  results.results_bundle(c["EBN0"], c["ESN"], c["MMSE"], meta=meta))
        frames = len(ebno) * a.blocks * a.frames_per_block
        print(json.dumps({"ebno": c["EBN0"], **{k: [round(v, 5) for v in c[k]] for k in c if not k.startswith("_") and k != "EBN0"},
                          "frames": frames, "seconds": round(dt, 2), "gpus": world}))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
