#!/usr/bin/env python
"""ESN hyper-parameter sweep (BASELINE.json configs[3]; the reference publishes the plots under
results/ESN_sweep_parameters/ but no generating script): reservoir size x spectral radius x sparsity,
uncoded BER vs Eb/N0 on the 4x8 block-fading 16-QAM link, every configuration its own reservoir.

  python examples/esn_sweep.py --out sweep_run                                  # one GPU
  torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 examples/esn_sweep.py --out ...

Whole configurations are placed on ranks (largest first, esn_b200.dist.assign_by_cost); there is no
data-path collective -- one allreduce at the end gathers the error counters.  The host-side weight
initialisation (pyESN.initweights order, libs/pyESN.py:93-109; the eigenvalue solve costs 3.5 s at
2048 neurons) of the next configuration runs on a host thread while the GPU works on the current one.
Every size detects on the tensor cores: up to 512 neurons on the kernel that keeps the state in shared memory,
larger reservoirs (600, 1024, 2048) on the streamed-state kernel (esn_predict_tcs).
"""
import argparse
import concurrent.futures as cf
import csv
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "esn-ofdm-mimo_b200"))


def init_weights(seed, N, ni, no, rho, sparsity):
    rng = np.random.RandomState(seed)
    W = rng.rand(N, N) - 0.5
    W[rng.rand(N, N) < sparsity] = 0
    W *= rho / np.max(np.abs(np.linalg.eigvals(W)))
    return W, rng.rand(N, ni) * 2 - 1, rng.rand(N, no) * 2 - 1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nt", type=int, default=4)
    ap.add_argument("--nr", type=int, default=8)
    ap.add_argument("--nsub", type=int, default=512)
    ap.add_argument("--qam-bits", type=int, default=4)
    ap.add_argument("--nres", default="64,128,256,512,1024,2048")
    ap.add_argument("--rho", default="0.7,0.9,1.1")
    ap.add_argument("--sparsity", default="0.1,0.3,0.5")
    ap.add_argument("--ebno", default="0:6:30", help="start:step:stop (inclusive) in dB")
    ap.add_argument("--blocks", type=int, default=16, help="coherence blocks per configuration and Eb/N0")
    ap.add_argument("--frames-per-block", type=int, default=128)
    ap.add_argument("--fit-precision", default="fp64", choices=["fp64", "fp32"])
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--out", default="esn_sweep_run")
    a = ap.parse_args()

    import torch
    import esn_b200
    from esn_b200 import Reservoir, dist as D, linksim
    rank, world, local = D.init_from_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    esn_b200.load()
    lo, st, hi = (float(x) for x in a.ebno.split(":"))
    ebno = list(np.arange(lo, hi + 1e-9, st))
    ni, no = 2 * a.nr, 2 * a.nt
    cfgs = [(int(n), float(r), float(s)) for n in a.nres.split(",") for r in a.rho.split(",") for s in a.sparsity.split(",")]
    # cost model: the recurrence, 2 N (N + n_in + n_out) flop per step; the streaming kernel (> 512) is ~12x slower per flop
    cost = [n * (n + ni + no) * (12.0 if n > 512 else 1.0) for n, _, _ in cfgs]
    mine = D.assign_by_cost(cost, world)[rank]
    counts = torch.zeros((len(cfgs), len(ebno), len(linksim.DETECTORS), 2), dtype=torch.int64, device=dev)
    secs = torch.zeros((len(cfgs),), dtype=torch.float64, device=dev)
    D.allreduce_sum_(secs)                                # NCCL communicator set-up stays out of the timing
    torch.cuda.synchronize()
    t_all = time.time()
    with cf.ThreadPoolExecutor(max_workers=2) as pool:
        futs = {i: pool.submit(init_weights, a.seed + i, cfgs[i][0], ni, no, cfgs[i][1], cfgs[i][2]) for i in mine}
        for i in mine:
            N, rho, sp = cfgs[i]
            W, W_in, W_fb = futs.pop(i).result()

            base = Reservoir(W, W_in, W_fb, 0.005 * np.ones(ni), np.zeros(ni), 5e-7 * np.ones(no), np.zeros(no), 0.001, True)

            def factory(var_x, base=base):                # same device weights, inputs rescaled per Eb/N0
                return base.rescaled(input_scaling=(0.005 / var_x ** 0.5) * np.ones(ni))
            t0 = time.time()
            c = linksim.ber_curve(factory, a.nt, a.nr, a.nsub, a.qam_bits, ebno, a.blocks, a.frames_per_block,
                                  seed=a.seed, fit_precision=a.fit_precision,
                                  detect_precision="tc", shard=False)
            torch.cuda.synchronize()
            counts[i] = c["_counts"]
            secs[i] = time.time() - t0
    D.allreduce_sum_(counts)
    D.allreduce_sum_(secs)
    wall = D.max_over_ranks(time.time() - t_all, dev)
    if rank == 0:
        os.makedirs(a.out, exist_ok=True)
        ber = (counts[..., 0].double() / counts[..., 1].clamp(min=1).double()).cpu().numpy()
        with open(os.path.join(a.out, "esn_sweep.csv"), "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(["n_reservoir", "spectral_radius", "sparsity", "detector", "seconds"] + ["EbNo_%gdB" % e for e in ebno])
            for i, (N, rho, sp) in enumerate(cfgs):
                for di, k in enumerate(linksim.DETECTORS):
                    if k == "ESN" or i == 0:             # the baselines do not depend on the reservoir
                        w.writerow([N, rho, sp, k, "%.2f" % float(secs[i])] + ["%.6f" % v for v in ber[i, :, di]])
        frames = len(cfgs) * len(ebno) * a.blocks * a.frames_per_block
        best = int(np.argmin(ber[:, -1, 0]))
        print(json.dumps({"configs": len(cfgs), "frames": frames, "seconds": round(wall, 2), "gpus": world,
                          "per_size_seconds": {str(n): round(float(sum(secs[i] for i, c in enumerate(cfgs) if c[0] == n)), 2)
                                               for n in sorted({c[0] for c in cfgs})},
                          "best_at_top_snr": {"n_reservoir": cfgs[best][0], "spectral_radius": cfgs[best][1],
                                              "sparsity": cfgs[best][2], "ber": round(float(ber[best, -1, 0]), 5)}}))
    D.barrier()
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
