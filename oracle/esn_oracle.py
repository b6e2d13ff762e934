"""CPU oracle for the ESN symbol-detection hot path (TEST INFRASTRUCTURE ONLY).

This module is a float64 numpy restatement of the algorithm the reference
(`aoschu/esn-ofdm-mimo`, `/root/reference`) runs on its hot path.  It exists to
*check* the CUDA engine; it is never on the product path.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs may import it.

Parity pin: the reference holds no tests or golden vectors of its own
(SURVEY.md §4), so the oracle is pinned against the *live reference* imported in
the build container: `tests/golden/make_golden.py` runs `/root/reference/libs`
on fixed seeds and stores its outputs under `tests/golden/*.npz`;
`tests/test_oracle_golden.py` asserts this restatement reproduces them.

Every function cites the reference lines it follows (paths relative to
`/root/reference/`).  Differences from the reference are structural only: the
per-step state noise is passed in as an explicit `[steps, N_res]` tensor of
uniforms (the reference draws `random_state_.rand(N_res)` inside `_update`,
`libs/pyESN.py:125`), so that the same tensor can be shipped to the GPU.
"""
from __future__ import annotations

import math
import numpy as np


# ----------------------------------------------------------------------------
# a1/a2: constructor-side helpers and reservoir initialisation
# ----------------------------------------------------------------------------

def broadcast_arg(s, n):
    """`correct_dimensions` (libs/pyESN.py:4-24): None stays None, scalar is
    repeated n times, 1-D must have length n, anything else is an error."""
    if s is None:
        return None
    s = np.array(s)
    if s.ndim == 0:
        return np.array([s] * n)
    if s.ndim == 1:
        if len(s) != n:
            raise ValueError("arg must have length " + str(n))
        return s
    raise ValueError("Invalid argument")


def resolve_rng(random_state):
    """RNG choice of `ESN.__init__` (libs/pyESN.py:79-87): a RandomState is
    used as is, a truthy seed builds one, anything falsy (None, 0) means numpy's
    global generator."""
    if isinstance(random_state, np.random.RandomState):
        return random_state
    if random_state:
        try:
            return np.random.RandomState(random_state)
        except TypeError as e:
            raise Exception("Invalid seed: " + str(e))
    return np.random.mtrand._rand


def init_weights(rng, n_inputs, n_outputs, n_reservoir, spectral_radius, sparsity):
    """`ESN.initweights` (libs/pyESN.py:93-109).  The draw ORDER is part of the
    contract: rand(N,N) for W, rand(N,N) for the sparsity mask, then after the
    eigenvalue rescale rand(N,n_in) and rand(N,n_out)."""
    N = n_reservoir
    W = rng.rand(N, N) - 0.5
    W[rng.rand(N, N) < sparsity] = 0
    radius = np.max(np.abs(np.linalg.eigvals(W)))
    W = W * (spectral_radius / radius)
    W_in = rng.rand(N, n_inputs) * 2 - 1
    W_fb = rng.rand(N, n_outputs) * 2 - 1
    return W, W_in, W_fb


# ----------------------------------------------------------------------------
# a4/a5: affine I/O maps
# ----------------------------------------------------------------------------

def scale_inputs(u, input_scaling, input_shift):
    """`_scale_inputs` (libs/pyESN.py:127-135).  The reference multiplies by
    `np.diag(scaling)` with a full dgemm; the zero terms add exactly 0.0, so a
    column-wise product is bit-identical for finite inputs."""
    if input_scaling is not None:
        u = u * np.asarray(input_scaling, dtype=float)[None, :]
    if input_shift is not None:
        u = u + input_shift
    return u


def scale_teacher(y, teacher_scaling, teacher_shift):
    """`_scale_teacher` (libs/pyESN.py:137-144)."""
    if teacher_scaling is not None:
        y = y * teacher_scaling
    if teacher_shift is not None:
        y = y + teacher_shift
    return y


def unscale_teacher(y, teacher_scaling, teacher_shift):
    """`_unscale_teacher` (libs/pyESN.py:146-152)."""
    if teacher_shift is not None:
        y = y - teacher_shift
    if teacher_scaling is not None:
        y = y / teacher_scaling
    return y


# ----------------------------------------------------------------------------
# a3: one reservoir step
# ----------------------------------------------------------------------------

def update(W, W_in, W_fb, x, u, y, noise, uniforms, teacher_forcing=True):
    """`ESN._update` (libs/pyESN.py:111-125):
    tanh(W x + W_in u [+ W_fb y]) + noise * (U[0,1) - 0.5)."""
    pre = W @ x + W_in @ u
    if teacher_forcing:
        pre = pre + W_fb @ y
    return np.tanh(pre) + noise * (uniforms - 0.5)


# ----------------------------------------------------------------------------
# a6: fit  (harvest + pinv readout + train-set prediction)
# ----------------------------------------------------------------------------

def harvest(W, W_in, W_fb, in_s, teach_s, noise, uniforms, teacher_forcing=True):
    """Harvest loop of `ESN.fit` (libs/pyESN.py:179-182): states[0] = 0 and row
    n >= 1 is driven by in_s[n] and the PREVIOUS teacher row.  `uniforms` is
    `[T-1, N_res]`; row n-1 is consumed at step n."""
    T = in_s.shape[0]
    states = np.zeros((T, W.shape[0]))
    for n in range(1, T):
        states[n] = update(W, W_in, W_fb, states[n - 1], in_s[n], teach_s[n - 1],
                           noise, uniforms[n - 1], teacher_forcing)
    return states


def fit(W, W_in, W_fb, inputs, outputs, transient, noise, uniforms,
        input_scaling=None, input_shift=None, teacher_scaling=None,
        teacher_shift=None, teacher_forcing=True):
    """`ESN.fit` (libs/pyESN.py:154-216) with identity output activation.

    Returns a dict with W_out [n_out, N_res+n_in], the train-set prediction on
    ALL T rows (libs/pyESN.py:212-216), the harvested states and the three
    `last*` vectors the reference remembers (libs/pyESN.py:195-197; note
    `lastinput` is the UNSCALED last input row)."""
    if inputs.ndim < 2:
        inputs = np.reshape(inputs, (len(inputs), -1))
    if outputs.ndim < 2:
        outputs = np.reshape(outputs, (len(outputs), -1))
    in_s = scale_inputs(inputs, input_scaling, input_shift)
    teach_s = scale_teacher(outputs, teacher_scaling, teacher_shift)
    states = harvest(W, W_in, W_fb, in_s, teach_s, noise, uniforms, teacher_forcing)
    ext = np.hstack((states, in_s))
    W_out = np.dot(np.linalg.pinv(ext[transient:, :]), teach_s[transient:, :]).T
    pred = unscale_teacher(np.dot(ext, W_out.T), teacher_scaling, teacher_shift)
    return dict(W_out=W_out, pred_train=pred, states=states,
                laststate=states[-1, :], lastinput=inputs[-1, :],
                lastoutput=teach_s[-1, :], in_s=in_s, teach_s=teach_s)


# ----------------------------------------------------------------------------
# a7: predict (free-running, output fed back)
# ----------------------------------------------------------------------------

def predict(W, W_in, W_fb, W_out, inputs, transient, noise, uniforms,
            x0=None, y0=None, input_scaling=None, input_shift=None,
            teacher_scaling=None, teacher_shift=None, teacher_forcing=True,
            return_states=False):
    """`ESN.predict` (libs/pyESN.py:218-255).  `x0`/`y0` are `laststate` /
    `lastoutput` when `continuation=True`, else None (zeros).  All T input rows
    drive the reservoir (unlike fit); `uniforms` is `[T, N_res]`.
    Returns rows `[transient:]` of the unscaled outputs."""
    if inputs.ndim < 2:
        inputs = np.reshape(inputs, (len(inputs), -1))
    T = inputs.shape[0]
    N = W.shape[0]
    n_out = W_out.shape[0]
    in_s = scale_inputs(inputs, input_scaling, input_shift)
    x = np.zeros(N) if x0 is None else np.asarray(x0, dtype=float)
    y = np.zeros(n_out) if y0 is None else np.asarray(y0, dtype=float)
    states = np.zeros((T, N))
    outs = np.zeros((T, n_out))
    for n in range(T):
        x = update(W, W_in, W_fb, x, in_s[n], y, noise, uniforms[n], teacher_forcing)
        y = W_out @ np.concatenate([x, in_s[n]])
        states[n] = x
        outs[n] = y
    res = unscale_teacher(outs[transient:], teacher_scaling, teacher_shift)
    if return_states:
        return res, states
    return res


# ----------------------------------------------------------------------------
# convenience: an object with the reference's call shape, explicit noise stream
# ----------------------------------------------------------------------------

class OracleESN:
    """Object wrapper with the reference's constructor/fit/predict call shape
    (libs/pyESN.py:31-255), consuming `random_state_` in the reference's order
    (init draws, then T-1 rows per fit, T rows per predict)."""

    def __init__(self, n_inputs, n_outputs, n_reservoir=200, spectral_radius=0.95,
                 sparsity=0, noise=0.001, input_shift=None, input_scaling=None,
                 teacher_forcing=True, feedback_scaling=None, teacher_scaling=None,
                 teacher_shift=None, random_state=None, silent=True):
        self.n_inputs, self.n_outputs, self.n_reservoir = n_inputs, n_outputs, n_reservoir
        self.spectral_radius, self.sparsity, self.noise = spectral_radius, sparsity, noise
        self.input_shift = broadcast_arg(input_shift, n_inputs)
        self.input_scaling = broadcast_arg(input_scaling, n_inputs)
        self.teacher_scaling, self.teacher_shift = teacher_scaling, teacher_shift
        self.teacher_forcing = teacher_forcing
        self.random_state_ = resolve_rng(random_state)
        self.W, self.W_in, self.W_feedb = init_weights(
            self.random_state_, n_inputs, n_outputs, n_reservoir, spectral_radius, sparsity)

    def _kw(self):
        return dict(input_scaling=self.input_scaling, input_shift=self.input_shift,
                    teacher_scaling=self.teacher_scaling, teacher_shift=self.teacher_shift,
                    teacher_forcing=self.teacher_forcing)

    def fit(self, inputs, outputs, transient=0):
        T = inputs.shape[0]
        uni = self.random_state_.rand(max(T - 1, 0), self.n_reservoir)
        r = fit(self.W, self.W_in, self.W_feedb, inputs, outputs, transient,
                self.noise, uni, **self._kw())
        self.W_out = r["W_out"]
        self.laststate, self.lastinput, self.lastoutput = r["laststate"], r["lastinput"], r["lastoutput"]
        self.last_fit = r
        return r["pred_train"]

    def predict(self, inputs, transient=0, continuation=True):
        T = inputs.shape[0]
        uni = self.random_state_.rand(T, self.n_reservoir)
        x0 = self.laststate if continuation else None
        y0 = self.lastoutput if continuation else None
        return predict(self.W, self.W_in, self.W_feedb, self.W_out, inputs, transient,
                       self.noise, uni, x0=x0, y0=y0, **self._kw())


# ----------------------------------------------------------------------------
# a8/a9: trainers
# ----------------------------------------------------------------------------

def pack_io(y_CP, x_CP, d, N, cp_len, N_t, N_r):
    """`build_io_for_delay` (libs/helper_mimo_esn_generic.py:26-38): re/im
    interleaved per antenna, d trailing zero rows on the input, teacher shifted
    down by d rows."""
    T = N + cp_len
    X_in = np.zeros((T + d, 2 * N_r))
    X_in[:T, 0::2] = y_CP[:, :N_r].real
    X_in[:T, 1::2] = y_CP[:, :N_r].imag
    X_out = np.zeros((T + d, 2 * N_t))
    X_out[d:d + T, 0::2] = x_CP[:, :N_t].real
    X_out[d:d + T, 1::2] = x_CP[:, :N_t].imag
    return X_in, X_out


def nmse_after_delay(pred, x_CP, d, N, N_t, isi_duration):
    """NMSE scoring of `nmse_for_delay` (libs/helper_mimo_esn_generic.py:47-56),
    including the quirk that `pred` (already stripped of d+CP rows) is sliced
    `[d:d+N+1]` again."""
    total = 0.0
    for tx in range(N_t):
        x_hat = pred[d:d + N + 1, 2 * tx] + 1j * pred[d:d + N + 1, 2 * tx + 1]
        x_true = x_CP[isi_duration - 1:, tx]
        M = min(len(x_hat), len(x_true))
        if M > 0:
            total += (np.linalg.norm(x_hat[:M] - x_true[:M]) ** 2
                      / (np.linalg.norm(x_true[:M]) ** 2 + 1e-12))
    return total


def train_generic(esn, DelayFlag, Min_Delay, Max_Delay, cp_len, N, N_t, N_r,
                  isi_duration, y_CP, x_CP):
    """`trainMIMOESN_generic` (libs/helper_mimo_esn_generic.py:5-86) on any
    object with the reference's fit/predict call shape."""
    def score(d):
        X_in, X_out = pack_io(y_CP, x_CP, d, N, cp_len, N_t, N_r)
        n_forget = d + cp_len
        esn.fit(X_in, X_out, n_forget)
        pred = esn.predict(X_in, n_forget, continuation=False)
        return nmse_after_delay(pred, x_CP, d, N, N_t, isi_duration), X_in, X_out, n_forget

    if DelayFlag == 0:
        d = int((Min_Delay + Max_Delay) // 2)
        nmse, X_in, X_out, n_forget = score(d)
        delay_idx, nmse_best = d - Min_Delay, float(nmse)
    else:
        nmse_best, best, delay_idx = 1e9, None, 0
        for dd in range(Min_Delay, Max_Delay + 1):
            nmse, a, b, nf = score(dd)
            if nmse < nmse_best:
                nmse_best, best, delay_idx = nmse, (a, b, nf, dd), dd - Min_Delay
        X_in, X_out, n_forget, d = best
        nmse_best = float(nmse_best)
    esn.fit(X_in, X_out, n_forget)
    Delay = np.full(2 * N_t, int(d), dtype=int)
    return [X_in, X_out, esn, Delay, delay_idx, int(d), int(d), n_forget, nmse_best]


def train_legacy_2x2(esn, DelayFlag, Min_Delay, Max_Delay, cp_len, N, N_t, N_r,
                     isi_duration, y_CP, x_CP, verbose=False):
    """`HelpFunc.trainMIMOESN` with DelayFlag == 0 (libs/HelpFunc.py:64-187):
    scans the shared delays 0..Max_Delay, then IGNORES the argmin and trains at
    row 3 (libs/HelpFunc.py:157-159).  Hard-wired to 2 Rx / 2 Tx columns.  The
    DelayFlag != 0 branch of the reference raises TypeError
    (libs/HelpFunc.py:76) and is reproduced as such."""
    if DelayFlag:
        raise TypeError("Cannot interpret '1' as a data type")
    lut = np.zeros(((Max_Delay + 1 - Min_Delay), 4)).astype('int32')
    for j in range(0, Max_Delay + 1):
        lut[j, :] = j
    dmax, dmin = np.amax(lut, axis=1), np.amin(lut, axis=1)

    def build(row):
        cur = lut[row]
        T = N + dmax[row] + cp_len
        ein, eout = np.zeros((T, N_t * 2)), np.zeros((T, N_t * 2))
        for c in range(2):
            ein[:, 2 * c] = np.append(y_CP[:, c].real, np.zeros(dmax[row]))
            ein[:, 2 * c + 1] = np.append(y_CP[:, c].imag, np.zeros(dmax[row]))
            eout[cur[2 * c]:cur[2 * c] + N + cp_len, 2 * c] = x_CP[:, c].real
            eout[cur[2 * c + 1]:cur[2 * c + 1] + N + cp_len, 2 * c + 1] = x_CP[:, c].imag
        return ein, eout

    nmse = np.zeros(lut.shape[0])
    for row in range(lut.shape[0]):
        cur = lut[row]
        ein, eout = build(row)
        n_forget = dmin[row] + cp_len
        esn.fit(ein, eout, n_forget)
        p = esn.predict(ein, n_forget, continuation=False)
        x = x_CP[isi_duration - 1:, :]
        for c in range(2):
            o = cur[2 * c] - dmin[row]
            oi = cur[2 * c + 1] - dmin[row]
            xh = p[o:o + N + 1, 2 * c] + 1j * p[oi:oi + N + 1, 2 * c + 1]
            nmse[row] += np.linalg.norm(xh - x[:, c]) ** 2 / np.linalg.norm(x[:, c]) ** 2
    delay_idx = 3
    if verbose:
        print(nmse)
    ein, eout = build(delay_idx)
    n_forget = dmin[delay_idx] + cp_len
    esn.fit(ein, eout, n_forget)
    return [ein, eout, esn, lut[delay_idx, :], delay_idx, dmin[delay_idx],
            dmax[delay_idx], n_forget, np.amin(nmse)]


# ----------------------------------------------------------------------------
# a10/a13: constellation, slicer, bit labels
# ----------------------------------------------------------------------------

def unit_qam_constellation(Bi):
    """`HelpFunc.UnitQamConstellation` (libs/HelpFunc.py:6-39): square QAM of
    unit mean power; point index = PamM * i_re + i_im (imaginary part varies
    fastest)."""
    pam_m = math.ceil(math.sqrt(2 ** Bi) / 2) * 2
    pam = np.arange(-(pam_m - 1), pam_m, 2).astype(float)
    C = (pam[:, None] + 1j * pam[None, :]).reshape(-1)
    return C / math.sqrt(np.mean(np.abs(C) ** 2))


def hard_demap_indices(X_hat, const):
    """Nearest-point search of `hard_bits_from_syms`
    (system_model_2/OFDM_MIMO_2-2_NBF_LDPC.py:103-111): argmin_k |const_k - s|,
    first minimum wins."""
    d = np.abs(const.reshape(1, -1) - np.asarray(X_hat).reshape(-1, 1))
    return np.argmin(d, axis=1).reshape(np.shape(X_hat))


def indices_to_bits(idx, m):
    """`bits_to_grayvec` (OFDM_MIMO_2-2_NBF_LDPC.py:36-38): plain binary of the
    point index, LSB first.  idx [N, N_t] -> bits [N*m, N_t]."""
    idx = np.asarray(idx)
    N, N_t = idx.shape
    bits = ((idx[:, None, :] >> np.arange(m)[None, :, None]) & 1)
    return bits.reshape(N * m, N_t)


def bits_to_indices(bits, m):
    """Tx mapping idx = sum_i bit_i 2^i (OFDM_MIMO_2-2_NBF_LDPC.py:402-404)."""
    Nm, N_t = bits.shape
    b = bits.reshape(Nm // m, m, N_t)
    return (b * (1 << np.arange(m))[None, :, None]).sum(axis=1)


def slicer_indices(X_hat, m):
    """Closed-form slicer equivalent to `hard_demap_indices` away from decision
    boundaries: per axis level = clamp(round((v*s+L-1)/2)),
    idx = L*i_re + i_im."""
    L = int(round(math.sqrt(2 ** m)))
    s = math.sqrt(2.0 * (L * L - 1) / 3.0)
    X_hat = np.asarray(X_hat)
    ire = np.clip(np.floor((X_hat.real * s + L) / 2.0), 0, L - 1).astype(int)
    iim = np.clip(np.floor((X_hat.imag * s + L) / 2.0), 0, L - 1).astype(int)
    return L * ire + iim


def boundary_distance(X_hat, m):
    """Distance of each symbol to the nearest slicer boundary (for the
    "within 1e-5 of a decision boundary" count of BASELINE.json)."""
    L = int(round(math.sqrt(2 ** m)))
    s = math.sqrt(2.0 * (L * L - 1) / 3.0)
    X_hat = np.asarray(X_hat)
    bnd = (np.arange(1, L) * 2 - L) / s

    def dist(v):
        return np.min(np.abs(v[..., None] - bnd), axis=-1)
    return np.minimum(dist(X_hat.real), dist(X_hat.imag))


# ----------------------------------------------------------------------------
# §8f row 3: soft outputs -- noise-variance estimate, max-log LLRs, logistic calibration
# ----------------------------------------------------------------------------

def qam_bit_labels(m):
    """`qam_bit_labels` (system_model_2/Demo_MIMO_4x8_Sionna_CDL_ESN_v2.py:60-64):
    labels[idx, b] = bit b of idx, LSB first (`bits_to_grayvec`, :30-32)."""
    idx = np.arange(2 ** m)
    return ((idx[:, None] >> np.arange(m)[None, :]) & 1).astype(int)


def sigma2_from_decision(X_hat_col, const):
    """`est_sigma2_from_decision` (Demo_MIMO_4x8_Sionna_CDL_ESN_v2.py:84-88): mean
    squared distance to the nearest constellation point, + 1e-12."""
    z = np.asarray(X_hat_col).reshape(-1, 1)
    idx = np.argmin(np.abs(z - const.reshape(1, -1)) ** 2, axis=1)
    return float(np.mean(np.abs(z[:, 0] - const[idx]) ** 2) + 1e-12)


def frame_sigma2(X_hat, const):
    """Per-frame noise variance: mean of the per-Tx estimates (:459)."""
    return float(np.mean([sigma2_from_decision(X_hat[:, tx], const) for tx in range(X_hat.shape[1])]))


def llrs_maxlog(z, const, labels, sigma2):
    """`qam_llrs_maxlog` (Demo_MIMO_4x8_Sionna_CDL_ESN_v2.py:66-82):
    LLR_b = (min_{c: bit b = 1} |z-c|^2 - min_{c: bit b = 0} |z-c|^2) / max(sigma2, 1e-12);
    positive = bit 0 more likely.  z [N] -> [N, m]."""
    d = np.abs(np.asarray(z).reshape(-1, 1) - const.reshape(1, -1)) ** 2
    m = labels.shape[1]
    out = np.zeros((d.shape[0], m))
    for b in range(m):
        out[:, b] = (d[:, labels[:, b] == 1].min(axis=1) - d[:, labels[:, b] == 0].min(axis=1)) / max(sigma2, 1e-12)
    return out


def frame_llrs(X_hat, m):
    """LLRs of one frame in the reference's layout (N, m, N_t) (:460-463) and its sigma2."""
    const, labels = unit_qam_constellation(m), qam_bit_labels(m)
    s2 = frame_sigma2(X_hat, const)
    return np.stack([llrs_maxlog(X_hat[:, tx], const, labels, s2) for tx in range(X_hat.shape[1])], axis=2), s2


def fit_logreg_1d(x, y, maxiter=400, lr=0.15, l2=1e-3):
    """`fit_logreg_1d` (Demo_MIMO_4x8_Sionna_CDL_ESN_v2.py:108-119): full-batch gradient
    descent on p(y=1|x) = sigmoid(a x + b), a0 = 1, b0 = 0."""
    a, b = 1.0, 0.0
    n = len(x)
    for _ in range(maxiter):
        with np.errstate(over="ignore"):
            p = 1.0 / (1.0 + np.exp(-(a * x + b)))
        ga = np.dot(p - y, x) / n + l2 * a
        gb = np.sum(p - y) / n
        a -= lr * ga
        b -= lr * gb
    return float(a), float(b)


def calibrate_llrs(llr, a, b, clip=20.0):
    """Calibrated, sign-flipped and clipped LLRs fed to the decoder (:489-492):
    clip(-(a_b llr + b_b), -clip, clip); llr [..., m, N_t]."""
    a = np.asarray(a).reshape(-1, 1)
    b = np.asarray(b).reshape(-1, 1)
    return np.clip(-(a * llr + b), -clip, clip)


# ----------------------------------------------------------------------------
# a11: ESN output unpack + FFT
# ----------------------------------------------------------------------------

def esn_output_to_freq(x_hat_tmp, N, N_t, Pi):
    """`reconstruct_esn_outputs_generic` with Delay == Delay_Min
    (OFDM_MIMO_2-2_NBF_LDPC.py:56-64) followed by (1/N) FFT / sqrt(Pi)
    (:435-438)."""
    X = np.zeros((N, N_t), dtype=complex)
    for tx in range(N_t):
        xt = x_hat_tmp[0:N + 1, 2 * tx] + 1j * x_hat_tmp[0:N + 1, 2 * tx + 1]
        X[:, tx] = (1.0 / N) * np.fft.fft(xt) / math.sqrt(Pi)
    return X


def pack_rx(y_CP, d):
    """Detect-time ESN input packing (OFDM_MIMO_2-2_NBF_LDPC.py:430-433)."""
    T, N_r = y_CP.shape
    X = np.zeros((T + d, 2 * N_r))
    X[:T, 0::2] = y_CP.real
    X[:T, 1::2] = y_CP.imag
    return X


# ----------------------------------------------------------------------------
# a12: baseline chain (Rx FFT, LS + interpolation + time-domain MMSE, ZF/MMSE)
# ----------------------------------------------------------------------------

def rx_fft(y_CP, cp_len, N):
    """Y = (1/N) FFT(y_CP[CP:]) (OFDM_MIMO_2-2_NBF_LDPC.py:428)."""
    return (1.0 / N) * np.fft.fft(y_CP[cp_len:, :], axis=0)


def interp_extrap_linear(xp, fp, x):
    """`interp1d(kind='linear', fill_value='extrapolate')`
    (OFDM_MIMO_2-2_NBF_LDPC.py:325-326) for sorted xp: linear inside, the end
    segments' lines outside."""
    xp = np.asarray(xp, dtype=float)
    j = np.clip(np.searchsorted(xp, x, side="right") - 1, 0, len(xp) - 2)
    t = (x - xp[j]) / (xp[j + 1] - xp[j])
    return fp[j] + t * (fp[j + 1] - fp[j])


def channel_estimate(Y_LS, X_LS, Pi, No, N, N_t, N_r, isi_magnitude, isi_duration):
    """Pilot LS on the comb `tx::N_t`, linear inter/extrapolation over all N
    subcarriers, IFFT, truncate to isi_duration taps, time-domain MMSE shrink,
    FFT back (OFDM_MIMO_2-2_NBF_LDPC.py:316-334).  Returns (H_LS, H_MMSE) of
    shape [N, N_r, N_t]."""
    H_LS = np.zeros((N, N_r, N_t), dtype=complex)
    H_MMSE = np.zeros((N, N_r, N_t), dtype=complex)
    R_h = np.diag(isi_magnitude[:isi_duration])
    mmse_scaler = (No / Pi) / (N / 2)
    A = np.dot(np.linalg.inv(R_h), mmse_scaler) + np.eye(isi_duration)
    k_all = np.arange(N)
    for nr in range(N_r):
        for tx in range(N_t):
            sc = np.arange(tx, N, N_t)
            h_sc = Y_LS[sc, nr] / (X_LS[sc, tx] * (Pi ** 0.5) + 1e-12)
            h_full = interp_extrap_linear(sc, h_sc, k_all)
            c_ls = np.fft.ifft(h_full)[:isi_duration]
            c_mmse = np.linalg.solve(A, c_ls)
            H_LS[:, nr, tx] = h_full
            H_MMSE[:, nr, tx] = np.fft.fft(np.r_[c_mmse, np.zeros(N - isi_duration)])
    return H_LS, H_MMSE


def equalize(Y, H, power_scale, reg):
    """`equalize_zf` (reg = 1e-12) / `equalize_mmse` (reg = No/Pi)
    (OFDM_MIMO_2-2_NBF_LDPC.py:41-53), applied per subcarrier as the loop at
    :453-460 does.  Y [N, N_r], H [N, N_r, N_t] -> X_hat [N, N_t]."""
    N, N_r, N_t = H.shape
    X = np.zeros((N, N_t), dtype=complex)
    for k in range(N):
        Hk = H[k]
        HH = Hk.conj().T
        G = HH @ Hk + reg * np.eye(N_t, dtype=complex)
        X[k] = np.linalg.solve(G, HH @ Y[k].reshape(N_r, 1)).reshape(-1) / power_scale
    return X


# ----------------------------------------------------------------------------
# workload generation (transmit chain + channel)
# ----------------------------------------------------------------------------

def isi_profile(isi_duration):
    """Exponential power-delay profile (OFDM_MIMO_2-2_NBF_LDPC.py:160-164)."""
    cp = isi_duration - 1
    mag = np.exp(-(np.arange(cp + 1)) / (cp / 9))
    return mag / np.sum(mag)


def draw_channel(rng, N_r, N_t, isi_magnitude, isi_duration):
    """Rayleigh block-fading taps (OFDM_MIMO_2-2_NBF_LDPC.py:272-280),
    c[nr, nt, tap]; draw order randn(re) then randn(im) per link."""
    c = np.zeros((N_r, N_t, isi_duration), dtype=complex)
    for nr in range(N_r):
        for nt in range(N_t):
            c0 = (rng.randn(isi_duration) + 1j * rng.randn(isi_duration)) / np.sqrt(2)
            c[nr, nt] = c0 * np.sqrt(isi_magnitude[:isi_duration])
    return c


# 3GPP TR 38.901 Table 7.7.2-2, TDL-B: normalised delays and powers [dB] (the table of the reference's CDL demo,
# system_model_2/Demo_MIMO_4x8_Sionna_CDL_ESN_v2.py:127-137)
TDLB_NORM_DELAYS = np.array([0.0000, 0.1072, 0.2155, 0.2095, 0.2870, 0.2986, 0.3752, 0.5055, 0.3681, 0.3697, 0.5700,
                             0.5283, 1.1021, 1.2756, 1.5474, 1.7842, 2.0169, 2.8294, 3.0219, 3.6187, 4.1067, 4.2790, 4.7834])
TDLB_POW_DB = np.array([0.0, -2.2, -4.0, -3.2, -9.8, -1.2, -3.4, -5.2, -7.6, -3.0, -8.9, -9.0, -4.8, -5.7, -7.5, -1.9,
                        -7.6, -12.2, -9.8, -11.4, -14.9, -9.2, -11.3])


def draw_channel_tdlb(rng, N_r, N_t, isi_duration, fs_hz=2 * 1.024e6, ds_ns=300.0):
    """`build_cdlb_mimo_taps` / `_gen_cdlb_impulse` (Demo_MIMO_4x8_Sionna_CDL_ESN_v2.py:139-177): per link, every
    TDL-B path gets a CN(0, p) gain, is split linearly between the two neighbouring sample taps, paths beyond
    `isi_duration` taps are dropped and the link is normalised to unit energy.  `rng`: a RandomState (the demo uses
    `default_rng`; the draw order per link -- re then im per path -- is the demo's)."""
    p = 10.0 ** (TDLB_POW_DB / 10.0)
    p = p / p.sum()
    d = TDLB_NORM_DELAYS * ds_ns * 1e-9 * fs_hz
    c = np.zeros((N_r, N_t, isi_duration), dtype=complex)
    for nr in range(N_r):
        for nt in range(N_t):
            h = np.zeros(isi_duration, dtype=complex)
            for k in range(len(d)):
                i0 = int(np.floor(d[k]))
                frac = d[k] - i0
                g = (rng.standard_normal() + 1j * rng.standard_normal()) / np.sqrt(2.0) * np.sqrt(p[k])
                if 0 <= i0 < isi_duration:
                    h[i0] += g * (1.0 - frac)
                if 0 <= i0 + 1 < isi_duration:
                    h[i0 + 1] += g * frac
            e = np.sum(np.abs(h) ** 2)
            c[nr, nt] = h / np.sqrt(e) if e > 0 else h
    return c


def fir_causal(h, x):
    """`scipy.signal.lfilter(h, [1], x)` for an FIR h: causal convolution
    truncated to len(x), zero initial state (OFDM_MIMO_2-2_NBF_LDPC.py:305)."""
    return np.convolve(x, h)[:len(x)]


def tx_frame(X, N, cp_len, Pi, A_clip):
    """Frequency symbols -> time domain with CP, power scaling and the soft PA
    clip with p_smooth = 1 (OFDM_MIMO_2-2_NBF_LDPC.py:413-419).  Returns
    (x_CP, x_CP_NLD), each [N+CP, N_t]."""
    x_t = N * np.fft.ifft(X, axis=0)
    x_cp = np.concatenate([x_t[-cp_len:], x_t], axis=0) * (Pi ** 0.5) if cp_len > 0 \
        else x_t * (Pi ** 0.5)
    x_nld = x_cp / ((1 + (np.abs(x_cp) / A_clip) ** 2) ** 0.5)
    return x_cp, x_nld


def channel_apply(rng, c, x_nld, No):
    """FIR channel per link plus AWGN of std sqrt((N+CP) No / 2) per component
    (OFDM_MIMO_2-2_NBF_LDPC.py:421-426)."""
    N_r, N_t, _ = c.shape
    T = x_nld.shape[0]
    y = np.zeros((T, N_r), dtype=complex)
    for nr in range(N_r):
        for tx in range(N_t):
            y[:, nr] += fir_causal(c[nr, tx], x_nld[:, tx])
        y[:, nr] += math.sqrt(T * No / 2) * (rng.randn(T) + 1j * rng.randn(T))
    return y


def synth_block(seed, N, N_t, N_r, m, ebno_db, n_data, isi_duration=8, No=1e-5,
                clip_db=3.0, channel="rayleigh"):
    """One coherence block of the block-fading template: a pilot frame plus
    `n_data` data frames through one channel draw (restates the generator of
    OFDM_MIMO_2-2_NBF_LDPC.py:270-312 and :387-426 with an explicit
    RandomState; the draw order differs from the script, which also interleaves
    ESN-internal draws on the global stream)."""
    rng = np.random.RandomState(seed)
    cp = isi_duration - 1
    const = unit_qam_constellation(m)
    Pi = 10 ** (ebno_db / 10) * No
    var_x = 10 ** (ebno_db / 10) * No * N
    A_clip = math.sqrt(var_x) * 10 ** (clip_db / 20)
    mag = isi_profile(isi_duration)
    c = draw_channel_tdlb(rng, N_r, N_t, isi_duration) if channel == "tdlb" else \
        draw_channel(rng, N_r, N_t, mag, isi_duration)
    H_true = np.fft.fft(np.concatenate(
        [c, np.zeros((N_r, N_t, N - isi_duration))], axis=2), axis=2).transpose(2, 0, 1)

    def one_frame():
        bits = (rng.rand(N * m, N_t) > 0.5).astype(np.int32)
        idx = bits_to_indices(bits, m)
        X = const[idx]
        x_cp, x_nld = tx_frame(X, N, cp, Pi, A_clip)
        return bits, idx, X, x_cp, x_nld

    bits_p, idx_p, X_p, x_cp_p, x_nld_p = one_frame()
    # comb pilot for LS (OFDM_MIMO_2-2_NBF_LDPC.py:287-289, :296-297, :301)
    X_LS = np.zeros_like(X_p)
    for tx in range(N_t):
        X_LS[tx::N_t, tx] = X_p[tx::N_t, tx]
    _, x_ls_nld = tx_frame(X_LS, N, cp, Pi, A_clip)
    T = N + cp
    y_p = np.zeros((T, N_r), dtype=complex)
    y_ls = np.zeros((T, N_r), dtype=complex)
    for nr in range(N_r):
        for tx in range(N_t):
            y_p[:, nr] += fir_causal(c[nr, tx], x_nld_p[:, tx])
            y_ls[:, nr] += fir_causal(c[nr, tx], x_ls_nld[:, tx])
        nz = math.sqrt(T * No / 2) * (rng.randn(T) + 1j * rng.randn(T))
        y_p[:, nr] += nz
        y_ls[:, nr] += nz
    data = []
    for _ in range(n_data):
        bits, idx, X, x_cp, x_nld = one_frame()
        y = channel_apply(rng, c, x_nld, No)
        data.append(dict(bits=bits, idx=idx, X=X, x_CP=x_cp, y_CP=y))
    return dict(c=c, H_true=H_true, Pi=Pi, No=No, var_x=var_x, cp=cp, const=const,
                isi_magnitude=mag,
                pilot=dict(bits=bits_p, idx=idx_p, X=X_p, X_LS=X_LS, x_CP=x_cp_p,
                           y_CP=y_p, y_LS_CP=y_ls),
                data=data)
