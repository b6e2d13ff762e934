#!/usr/bin/env python
"""Runs examples/siso_qpsk_awgn.py's demo loop with the LIVE reference pyESN (/root/reference/libs, CPU numpy)
and stores its error counts: tests/golden/siso_demo_golden.npz.  Run in the build container only."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, "/root/reference/libs")
sys.path.insert(0, os.path.join(ROOT, "examples"))
from pyESN import ESN  # noqa: E402  (the reference's)
import siso_qpsk_awgn as demo  # noqa: E402

EBNO, SYMBOLS, NRES, SEED = [3.0, 9.0, 15.0], 6, 200, 42
assert ESN.__module__ == "pyESN" and "/root/reference" in sys.modules["pyESN"].__file__
r = demo.run(ESN, EBNO, SYMBOLS, NRES, seed=SEED, keep_first=True)
np.savez_compressed(os.path.join(HERE, "siso_demo_golden.npz"), ebno=np.array(EBNO), symbols=SYMBOLS, nres=NRES,
                    seed=SEED, bits=np.array(r["bits"]), esn_per_symbol=np.array(r["esn_per_symbol"]),
                    first_xhat=np.array(r["first_xhat"]),
                    **{"err_" + k: np.array(r[k]) for k in ("ESN", "MMSE", "ZF", "LS")})
print({k: r[k] for k in ("ESN", "MMSE", "ZF", "LS", "bits", "esn_per_symbol", "near_boundary")})
