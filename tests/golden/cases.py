"""Deterministic input builders shared by `make_golden.py` (which runs the live
reference on them) and by the parity tests (which run the oracle / the CUDA
engine on the same inputs).  Nothing here touches /root/reference."""
from __future__ import annotations

import math
import numpy as np

# name -> ESN/problem shape.  `rows` = N_sub + CP + d (time steps per frame).
ESN_CASES = {
    # tiny SISO, default-transient fit, continuation predict (SISO demo call shape)
    "siso_small": dict(n_in=2, n_out=2, n_res=40, T=48, transient=0, sparsity=0.1,
                       rho=0.9, noise=0.001, seed=7, in_scale=0.05, t_scale=5e-3,
                       continuation=True, teacher_forcing=True),
    # 2x2-shaped, overdetermined (rows used 64 > cols 54)
    "mimo2x2_small": dict(n_in=4, n_out=4, n_res=50, T=74, transient=10, sparsity=0.1,
                          rho=0.9, noise=0.001, seed=11, in_scale=0.01, t_scale=5e-7,
                          continuation=False, teacher_forcing=True),
    # 4x8-shaped, UNDERdetermined (rows used 32 < cols 80) -> min-norm solution
    "mimo4x8_small": dict(n_in=16, n_out=8, n_res=64, T=42, transient=10, sparsity=0.1,
                          rho=0.9, noise=0.001, seed=13, in_scale=0.01, t_scale=5e-7,
                          continuation=False, teacher_forcing=True),
    # noise-free and no-feedback variants
    "noise0_small": dict(n_in=4, n_out=4, n_res=48, T=90, transient=10, sparsity=0.0,
                         rho=0.9, noise=0.0, seed=17, in_scale=0.01, t_scale=5e-7,
                         continuation=False, teacher_forcing=True),
    "nofeedback_small": dict(n_in=4, n_out=4, n_res=48, T=90, transient=5, sparsity=0.3,
                             rho=1.1, noise=0.001, seed=19, in_scale=0.02, t_scale=1e-6,
                             continuation=False, teacher_forcing=False),
    # BASELINE.json configs[1]: 2x2, 100 neurons, N_sub 512 (T = 512+7+3)
    "cfg2_2x2_n100": dict(n_in=4, n_out=4, n_res=100, T=522, transient=10, sparsity=0.0,
                          rho=0.9, noise=0.001, seed=42, in_scale=0.005, t_scale=5e-7,
                          continuation=False, teacher_forcing=True),
    # BASELINE.json configs[2]: 4x8, 512 neurons, N_sub 512
    "cfg3_4x8_n512": dict(n_in=16, n_out=8, n_res=512, T=522, transient=10, sparsity=0.1,
                          rho=0.9, noise=0.001, seed=42, in_scale=0.005, t_scale=5e-7,
                          continuation=False, teacher_forcing=True),
}


def esn_kwargs(c):
    return dict(n_inputs=c["n_in"], n_outputs=c["n_out"], n_reservoir=c["n_res"],
                spectral_radius=c["rho"], sparsity=c["sparsity"], noise=c["noise"],
                input_shift=np.zeros(c["n_in"]),
                input_scaling=c["in_scale"] * np.ones(c["n_in"]),
                teacher_scaling=c["t_scale"] * np.ones(c["n_out"]),
                teacher_shift=np.zeros(c["n_out"]),
                teacher_forcing=c["teacher_forcing"], random_state=c["seed"])


def esn_io(c, which=0):
    """Synthetic time-domain training pair with the statistics of the demos:
    unit-variance complex-ish inputs, teacher = a short FIR mix of the inputs
    plus a little noise (so the readout problem is well posed but not trivial).
    `which` selects independent realisations (0 = training, 1.. = test)."""
    rng = np.random.RandomState(1000 * c["seed"] + which)
    T, n_in, n_out = c["T"], c["n_in"], c["n_out"]
    u = rng.randn(T, n_in)
    mix = rng.randn(3, n_in, n_out) / math.sqrt(n_in)
    y = np.zeros((T, n_out))
    for k in range(3):
        y[k:] += u[:T - k] @ mix[k]
    y += 0.01 * rng.randn(T, n_out)
    # the demos' teachers are O(1)-O(1e-2); keep O(1)
    return u, y


# trainer-level cases (complex y_CP / x_CP of a coherence block)
TRAINER_CASES = {
    "gen_2x2": dict(N=32, N_t=2, N_r=2, n_res=40, isi=8, seed=23, ebno=12, m=4),
    "gen_4x8": dict(N=32, N_t=4, N_r=8, n_res=48, isi=8, seed=29, ebno=15, m=4),
    "gen_1x2": dict(N=32, N_t=1, N_r=2, n_res=32, isi=8, seed=31, ebno=9, m=4),
    # BASELINE.json configs[2] at full size: 4x8, N_sub = 512, 512 neurons (T = 522, underdetermined 512 x 528 fit)
    "gen_4x8_n512": dict(N=512, N_t=4, N_r=8, n_res=512, isi=8, seed=37, ebno=15, m=4),
}


def trainer_esn_kwargs(c, var_x):
    return dict(n_inputs=2 * c["N_r"], n_outputs=2 * c["N_t"], n_reservoir=c["n_res"],
                spectral_radius=0.9, sparsity=0.1,
                input_shift=np.zeros(2 * c["N_r"]),
                input_scaling=(0.005 / (var_x ** 0.5)) * np.ones(2 * c["N_r"]),
                teacher_scaling=5e-7 * np.ones(2 * c["N_t"]),
                teacher_shift=np.zeros(2 * c["N_t"]),
                feedback_scaling=np.zeros(2 * c["N_t"]), random_state=c["seed"])


# soft-output cases: qam_bits -> (N, N_t, SNR of the synthetic equaliser output in dB, frames)
SOFT_CASES = {2: (64, 2, 6.0, 6), 4: (64, 4, 14.0, 6), 6: (32, 2, 20.0, 4)}


def soft_frames(m, N, N_t, snr_db, frames, const):
    """Seeded equaliser outputs X_hat [frames, N, N_t] = const[idx] + CN(0, 10^(-snr/10)) and idx."""
    rng = np.random.RandomState(900 + m)
    idx = rng.randint(0, 2 ** m, size=(frames, N, N_t))
    s = 10 ** (-snr_db / 20) / np.sqrt(2)
    X = const[idx] + s * (rng.randn(frames, N, N_t) + 1j * rng.randn(frames, N, N_t))
    return X, idx
