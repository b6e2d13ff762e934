"""Torch-tensor wrappers of the OFDM-side kernels (csrc/ofdm.cu): Rx FFT, ESN
output unpack+FFT+demap, LS/MMSE channel estimation, ZF/MMSE equalisation and
hard-decision error counting.  Reference sites: system_model_2/
OFDM_MIMO_2-2_NBF_LDPC.py:41-64, :103-111, :316-334, :428-474."""
from __future__ import annotations

import torch

from . import _lib
from ._lib import check, ptr
from .engine import _CODE, _stream


def _cplx_view(t):
    """complex tensor -> real view with trailing dim 2 (no copy)."""
    return torch.view_as_real(t.contiguous())


def _real_dtype(t):
    return {torch.complex64: torch.float32, torch.complex128: torch.float64}.get(t.dtype, t.dtype)


def _dev(v, dtype, device):
    """Python / numpy scalars become a device tensor through a fill kernel: no pageable host-to-device copy,
    hence no stream synchronisation in the middle of an enqueued chain."""
    if isinstance(v, torch.Tensor):
        return v.to(device=device, dtype=dtype)
    if not hasattr(v, "__len__"):
        return torch.full((1,), float(v), dtype=dtype, device=device)
    return torch.as_tensor(v, dtype=dtype, device=device)


def _scalar_vec(v, B, dtype, device):
    t = _dev(v, dtype, device).reshape(-1)
    if t.numel() not in (1, B):
        raise ValueError("expected a scalar or one value per frame")
    return t.contiguous(), (0 if t.numel() == 1 else 1)


def unpack_fft_demap(y, N, N_t, Pi, qam_bits, tx_idx=None, want_xhat=True, want_idx=True,
                     boundary_eps=0.0, counts=None):
    """y [B, rows>=N, 2*N_t] real ESN outputs -> (X_hat [B,N,N_t] complex,
    idx [B,N,N_t] uint8, counts [2] int64 = [bit errors, near-boundary symbols])."""
    lib = _lib.load()
    y = y.contiguous()
    B, rows, _ = y.shape
    code = _CODE[y.dtype]
    pi, stride = _scalar_vec(Pi, B, y.dtype, y.device)
    cd = torch.complex64 if y.dtype == torch.float32 else torch.complex128
    X = torch.empty((B, N, N_t), dtype=cd, device=y.device) if want_xhat else None
    idx = torch.empty((B, N, N_t), dtype=torch.uint8, device=y.device) if want_idx else None
    if counts is None:
        counts = torch.zeros(2, dtype=torch.int64, device=y.device)
    if tx_idx is not None:
        tx_idx = tx_idx.to(device=y.device, dtype=torch.uint8).contiguous()
    check(lib.ofdm_unpack_fft_demap(code, ptr(y), B, rows, N, N_t, ptr(pi), stride, qam_bits,
                                    ptr(torch.view_as_real(X)) if X is not None else None, ptr(idx),
                                    ptr(tx_idx), float(boundary_eps), ptr(counts), _stream()),
          "ofdm_unpack_fft_demap")
    return X, idx, counts


def rx_fft(y_cp, N, cp):
    """y_cp [B, N+cp, N_r] complex -> Y [B, N, N_r] = (1/N) FFT(y_cp[cp:])."""
    lib = _lib.load()
    B, _, N_r = y_cp.shape
    v = _cplx_view(y_cp)
    Y = torch.empty((B, N, N_r), dtype=y_cp.dtype, device=y_cp.device)
    check(lib.ofdm_rx_fft(_CODE[v.dtype], ptr(v), B, N, cp, N_r, ptr(torch.view_as_real(Y)), _stream()),
          "ofdm_rx_fft")
    return Y


def chanest(Y_LS, X_LS, Pi, isi_magnitude, taps, No):
    """LS + interpolation + time-domain MMSE.  Y_LS [B,N,N_r], X_LS [B,N,N_t]
    complex -> (H_LS, H_MMSE) [B,N,N_r,N_t] complex."""
    lib = _lib.load()
    B, N, N_r = Y_LS.shape
    N_t = X_LS.shape[2]
    rd = _real_dtype(Y_LS)
    pi = _dev(Pi, rd, Y_LS.device).reshape(-1)
    if pi.numel() == 1:
        pi = pi.expand(B)
    pi = pi.contiguous()
    mag = _dev(isi_magnitude, rd, Y_LS.device).contiguous()
    H_LS = torch.empty((B, N, N_r, N_t), dtype=Y_LS.dtype, device=Y_LS.device)
    H_MM = torch.empty_like(H_LS)
    check(lib.ofdm_chanest(_CODE[rd], ptr(_cplx_view(Y_LS)), ptr(_cplx_view(X_LS)), B, N, N_r, N_t, ptr(pi),
                           ptr(mag), int(taps), float(No), ptr(torch.view_as_real(H_LS)),
                           ptr(torch.view_as_real(H_MM)), _stream()), "ofdm_chanest")
    return H_LS, H_MM


def equalize(Y, H, reg, power_scale, h_index=None):
    """X_hat[b,k] = solve(H^H H + reg I, H^H Y[b,k]) / power_scale.  Y [B,N,N_r],
    H [Bh,N,N_r,N_t] complex (h_index[b] picks the block estimate)."""
    lib = _lib.load()
    B, N, N_r = Y.shape
    N_t = H.shape[3]
    rd = _real_dtype(Y)
    r, rs = _scalar_vec(reg, B, rd, Y.device)
    p, pst = _scalar_vec(power_scale, B, rd, Y.device)
    if h_index is not None:
        h_index = h_index.to(device=Y.device, dtype=torch.int32).contiguous()
    elif H.shape[0] != B:
        raise ValueError("h_index required when H holds fewer estimates than frames")
    X = torch.empty((B, N, N_t), dtype=Y.dtype, device=Y.device)
    check(lib.ofdm_equalize(_CODE[rd], ptr(_cplx_view(Y)), ptr(_cplx_view(H)), ptr(h_index), B, N, N_r, N_t,
                            ptr(r), rs, ptr(p), pst, ptr(torch.view_as_real(X)), _stream()), "ofdm_equalize")
    return X


def demap_count(X_hat, qam_bits, tx_idx=None, boundary_eps=0.0, counts=None, want_idx=True):
    lib = _lib.load()
    B, N, N_t = X_hat.shape
    rd = _real_dtype(X_hat)
    idx = torch.empty((B, N, N_t), dtype=torch.uint8, device=X_hat.device) if want_idx else None
    if counts is None:
        counts = torch.zeros(2, dtype=torch.int64, device=X_hat.device)
    if tx_idx is not None:
        tx_idx = tx_idx.to(device=X_hat.device, dtype=torch.uint8).contiguous()
    check(lib.ofdm_demap_count(_CODE[rd], ptr(_cplx_view(X_hat)), B, N, N_t, qam_bits, ptr(idx), ptr(tx_idx),
                               float(boundary_eps), ptr(counts), _stream()), "ofdm_demap_count")
    return idx, counts


def soft_demap(X_hat, qam_bits, cal=None, clip=20.0, want_llr=True):
    """X_hat [B,N,N_t] complex -> (sigma2 [B], llr [B,N,qam_bits,N_t]): the frame's noise-variance
    estimate and max-log LLRs (positive = bit 0).  `cal` = ab [qam_bits,2] from `llr_calibrate`
    applies the decoder-side map clip(-(a llr + b), +-clip)."""
    lib = _lib.load()
    B, N, N_t = X_hat.shape
    rd = _real_dtype(X_hat)
    sigma2 = torch.empty((B,), dtype=rd, device=X_hat.device)
    llr = torch.empty((B, N, qam_bits, N_t), dtype=rd, device=X_hat.device) if want_llr else None
    a = b = None
    if cal is not None:
        cal = cal.to(device=X_hat.device, dtype=torch.float64)
        a, b = cal[:, 0].contiguous(), cal[:, 1].contiguous()
    check(lib.ofdm_soft_demap(_CODE[rd], ptr(_cplx_view(X_hat)), B, N, N_t, qam_bits, ptr(a), ptr(b), float(clip),
                              ptr(sigma2), ptr(llr), _stream()), "ofdm_soft_demap")
    return sigma2, llr


def llr_calibrate(llr, tx_idx, qam_bits, maxiter=400, lr=0.1, l2=1e-3):
    """Per-bit logistic calibration of LLRs [B,N,qam_bits,N_t] against the transmitted symbol
    indices [B,N,N_t]; returns ab [qam_bits,2] fp64 on the device."""
    lib = _lib.load()
    llr = llr.contiguous()
    B, N, m, N_t = llr.shape
    tx_idx = tx_idx.to(device=llr.device, dtype=torch.uint8).contiguous()
    ab = torch.empty((qam_bits, 2), dtype=torch.float64, device=llr.device)
    check(lib.ofdm_llr_calibrate(_CODE[llr.dtype], ptr(llr), ptr(tx_idx), B, N, N_t, qam_bits, int(maxiter),
                                 float(lr), float(l2), ptr(ab), _stream()), "ofdm_llr_calibrate")
    return ab


def synth_frames(tx_idx, taps, Pi, A_clip, N, cp, qam_bits, noise_std, delay=0, chan_index=None, noise=None,
                 seed=0, dtype=torch.float32, want_x_cp=False, want_y_cp=True, want_esn_in=True):
    """Workload generation on the device (bits/indices -> received frames).  tx_idx [B,N,N_t]
    uint8, taps [n_chan,N_r,N_t,ntaps] complex, Pi / A_clip scalars or [B].  Returns a dict with
    x_cp [B,N+cp,N_t], y_cp [B,N+cp,N_r] (complex) and esn_in [B,N+cp+delay,2N_r] (real)."""
    lib = _lib.load()
    dev = tx_idx.device
    B, _, N_t = tx_idx.shape
    cd = torch.complex64 if dtype == torch.float32 else torch.complex128
    taps = taps.to(device=dev, dtype=cd).contiguous()
    _, N_r, _, ntaps = taps.shape
    tx_idx = tx_idx.to(torch.uint8).contiguous()
    pi = _dev(Pi, dtype, dev).reshape(-1)
    pi = (pi.expand(B) if pi.numel() == 1 else pi).contiguous()
    ac = _dev(A_clip, dtype, dev).reshape(-1)
    ac = (ac.expand(B) if ac.numel() == 1 else ac).contiguous()
    if chan_index is not None:
        chan_index = chan_index.to(device=dev, dtype=torch.int32).contiguous()
    elif taps.shape[0] != B:
        raise ValueError("chan_index required when fewer channel realisations than frames are given")
    if noise is not None:
        noise = noise.to(device=dev, dtype=cd).contiguous()
    L = N + cp
    x_cp = torch.empty((B, L, N_t), dtype=cd, device=dev) if want_x_cp else None
    y_cp = torch.empty((B, L, N_r), dtype=cd, device=dev) if want_y_cp else None
    esn_in = torch.empty((B, L + delay, 2 * N_r), dtype=dtype, device=dev) if want_esn_in else None
    rv = lambda t: None if t is None else ptr(torch.view_as_real(t))    # noqa: E731
    check(lib.ofdm_synth_frames(_CODE[dtype], ptr(tx_idx), rv(taps), ptr(chan_index), ptr(pi), ptr(ac), rv(noise),
                                float(noise_std), int(seed) & 0xFFFFFFFFFFFFFFFF, B, N, cp, N_t, N_r, ntaps,
                                qam_bits, int(delay), rv(x_cp), rv(y_cp), ptr(esn_in), _stream()),
          "ofdm_synth_frames")
    return dict(x_cp=x_cp, y_cp=y_cp, esn_in=esn_in)
