#!/usr/bin/env python
"""ONE readout trained on pilots that are spread over the GPUs of a box (BASELINE.json configs[4]): every
rank harvests its own pilots and accumulates the fp64 normal equations G = E^T E, R = E^T D; one NCCL
allreduce sums them; every rank runs the same Cholesky.  The result is checked against the readout rank 0
obtains from ALL pilots on its own GPU (must agree to fp64 round-off).

  torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 examples/shared_readout_nccl.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "esn-ofdm-mimo_b200"))


def main():
    import torch
    import esn_b200
    from esn_b200 import Reservoir, dist as D
    rank, world, local = D.init_from_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    esn_b200.load()
    N, ni, no, T, tr, n_pil = 512, 16, 8, 522, 10, 64
    rng = np.random.RandomState(42)
    W = rng.rand(N, N) - 0.5
    W[rng.rand(N, N) < 0.1] = 0
    W *= 0.9 / np.max(np.abs(np.linalg.eigvals(W)))
    res = Reservoir(W, rng.rand(N, ni) * 2 - 1, rng.rand(N, no) * 2 - 1, 0.005 * np.ones(ni), np.zeros(ni),
                    5e-7 * np.ones(no), np.zeros(no), 0.001, True, device=dev)
    g = torch.Generator(device="cpu").manual_seed(7)            # the same pilots on every rank
    u = torch.randn((n_pil, T, ni), generator=g, dtype=torch.float64).to(dev)
    y = (torch.randn((n_pil, T, no), generator=g, dtype=torch.float64) * 1e-2).to(dev)
    uni = torch.rand((n_pil, T - 1, N), generator=g, dtype=torch.float64).to(dev)
    b0, b1 = D.shard_range(n_pil, rank, world)
    # chunked: the all-reduce of a chunk's partial normal equations (2.2 MB + 33 KB) runs behind the next chunk's harvest
    W_shared, info, nbytes = res.train_shared_readout(u[b0:b1], y[b0:b1], tr, precision="fp64", chunks=2,
                                                      noise_uniforms=uni[b0:b1])
    ok = int(info.abs().max()) == 0
    ext_all = res.harvest(u, y, precision="fp64", noise_uniforms=uni)
    W_single, info1 = res.train_readout(ext_all, y, tr, shared=True)
    err = float((W_shared - W_single).norm() / W_single.norm())
    errs = torch.tensor([err], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(errs, op=torch.distributed.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"gpus": world, "pilots": n_pil, "rel_err_vs_single_gpu": float(errs[0]), "cholesky_ok": ok, "allreduce_bytes": int(nbytes),
                          "backend": torch.distributed.get_backend() if world > 1 else "none"}))
    assert ok and float(errs[0]) < 1e-9
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
