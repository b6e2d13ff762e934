"""Drop-in `helper_mimo_esn_generic` (reference libs/helper_mimo_esn_generic.py).

`trainMIMOESN_generic` keeps the reference's signature and 9-element return
list.  The packing of complex pilots into real ESN I/O is vectorised numpy on
the host (a few KB); every fit/predict it triggers runs on the GPU through the
`esn` object (the drop-in `pyESN.ESN`)."""
import numpy as np


def _pack(y_CP, x_CP, d, T, N_t, N_r):
    """Real I/O for a shared output delay d: [re, im] interleaved per antenna,
    input zero-padded by d rows at the end, teacher shifted down by d rows
    (reference :26-38)."""
    X_in = np.zeros((T + d, 2 * N_r), dtype=float)
    X_in[:T, 0::2] = y_CP[:, :N_r].real
    X_in[:T, 1::2] = y_CP[:, :N_r].imag
    X_out = np.zeros((T + d, 2 * N_t), dtype=float)
    X_out[d:d + T, 0::2] = x_CP[:, :N_t].real
    X_out[d:d + T, 1::2] = x_CP[:, :N_t].imag
    return X_in, X_out


def _nmse(pred, d, N, N_t, IsiDuration, x_CP):
    """The reference's NMSE of a prediction at delay d (reference :47-56, including its
    habit of slicing the already-trimmed prediction at [d:d+N+1])."""
    x_true = x_CP[IsiDuration - 1:, :N_t]
    seg = pred[d:d + N + 1]
    M = min(seg.shape[0], x_true.shape[0])
    nmse = 0.0
    if M > 0:
        x_hat = seg[:M, 0::2] + 1j * seg[:M, 1::2]
        for tx in range(N_t):
            num = np.linalg.norm(x_hat[:, tx] - x_true[:M, tx]) ** 2
            nmse += num / (np.linalg.norm(x_true[:M, tx]) ** 2 + 1e-12)
    return nmse


def _score(esn, d, CyclicPrefixLen, N, N_t, N_r, IsiDuration, y_CP, x_CP):
    """fit + predict at delay d and its NMSE (reference :40-56)."""
    X_in, X_out = _pack(y_CP, x_CP, d, N + CyclicPrefixLen, N_t, N_r)
    nForget = d + CyclicPrefixLen
    esn.fit(X_in, X_out, nForget)
    pred = esn.predict(X_in, nForget, continuation=False)
    return _nmse(pred, d, N, N_t, IsiDuration, x_CP), X_in, X_out, nForget


def trainMIMOESN_generic(esn, DelayFlag, Min_Delay, Max_Delay,
                         CyclicPrefixLen, N, N_t, N_r, IsiDuration,
                         y_CP, x_CP):
    """Train `esn` on one pilot OFDM symbol.

    DelayFlag == 0: single shared delay (Min+Max)//2; otherwise scan
    [Min_Delay, Max_Delay] and keep the first delay with the lowest NMSE.
    Returns [ESN_input, ESN_output, esn, Delay, Delay_Idx, Delay_Minn,
    Delay_Maxx, nForgetPoints, NMSE_ESN] (reference :59-86)."""
    args = (CyclicPrefixLen, N, N_t, N_r, IsiDuration, y_CP, x_CP)
    if DelayFlag == 0:
        d = int((Min_Delay + Max_Delay) // 2)
        nmse, ESN_input, ESN_output, nForgetPoints = _score(esn, d, *args)
    else:
        nmse, chosen = 1e9, None
        cands = list(range(Min_Delay, Max_Delay + 1))
        if hasattr(esn, "fit_predict_many"):
            # every candidate delay in one batched harvest / solve / predict (same noise order, same results)
            io = [_pack(y_CP, x_CP, c, N + CyclicPrefixLen, N_t, N_r) for c in cands]
            preds = esn.fit_predict_many([(xi, xo, c + CyclicPrefixLen) for (xi, xo), c in zip(io, cands)])
            scored = [(_nmse(p, c, N, N_t, IsiDuration, x_CP), xi, xo, c + CyclicPrefixLen)
                      for p, (xi, xo), c in zip(preds, io, cands)]
        else:
            scored = [_score(esn, c, *args) for c in cands]
        for cand, (s, xi, xo, nf) in zip(cands, scored):
            if s < nmse:
                nmse, chosen = s, (xi, xo, nf, cand)
        ESN_input, ESN_output, nForgetPoints, d = chosen
        d = int(d)
    esn.fit(ESN_input, ESN_output, nForgetPoints)       # final fit on the chosen delay
    Delay = np.full(2 * N_t, d, dtype=int)
    return [ESN_input, ESN_output, esn, Delay, d - Min_Delay, d, d, nForgetPoints, float(nmse)]
