"""Drop-in `HelpFunc` (reference libs/HelpFunc.py): QAM table, the unused
channel-correlation helper, and the legacy 2x2 trainer.  The functions live in
a class namespace and are called unbound (`HelpFunc.trainMIMOESN(esn, ...)`),
as the reference's callers do (system_model_2/system_model_2_all_comparision.py:441)."""
import math

import numpy as np


class HelpFunc():

    def UnitQamConstellation(Bi):
        """Square QAM with unit mean power; point index = side*i_re + i_im, the
        imaginary part varying fastest (reference :6-39)."""
        side = math.ceil(math.sqrt(2 ** Bi) / 2) * 2
        pam = np.arange(-(side - 1), side, 2).astype(np.int32)
        grid = pam[:, None] + 1j * pam[None, :]            # [i_re, i_im]
        C = grid.reshape(-1).astype('complex128')
        return C / math.sqrt(np.mean(abs(C) ** 2))

    def ComputeChannelCorrMatrix(IsiMagnitude):
        """Frequency-domain channel correlation matrix R_f[n, m] = r_f(n - m)
        with r_f = FFT(power delay profile) and r_f(-k) = conj(r_f(k))
        (reference :41-62)."""
        r = np.fft.fft(IsiMagnitude)
        n = np.arange(len(IsiMagnitude))
        diff = n[:, None] - n[None, :]
        R = np.where(diff >= 0, r[np.abs(diff)], np.conjugate(r[np.abs(diff)]))
        return R.astype('complex128')

    def trainMIMOESN(esn, DelayFlag, Min_Delay, Max_Delay, CyclicPrefixLen, N, N_t, N_r, IsiDuration, y_CP, x_CP):
        """Legacy trainer, hard-wired to 2 Rx / 2 Tx streams (4 real inputs, 4
        real outputs).  DelayFlag == 0 scans the shared delays 0..Max_Delay,
        prints the NMSE vector, then trains at table row 3 regardless of the
        argmin -- all as the reference does (:98-187).  The DelayFlag != 0
        branch of the reference dies with a TypeError before training
        (`np.zeros(shape, 1)`, :76); that behaviour is kept."""
        if DelayFlag:
            n_rows = (Max_Delay + 1 - Min_Delay) ** 2
            np.zeros(n_rows, 1)                       # raises TypeError, as the reference
        table = np.zeros(((Max_Delay + 1 - Min_Delay), 4)).astype('int32')
        for row in range(0, Max_Delay + 1):
            table[row, :] = row
        hi, lo = np.amax(table, axis=1), np.amin(table, axis=1)
        T = N + CyclicPrefixLen

        def build(row):
            dl = table[row]
            ein = np.zeros((T + hi[row], N_t * 2))
            eout = np.zeros((T + hi[row], N_t * 2))
            for s in range(2):
                ein[:T, 2 * s] = y_CP[:, s].real
                ein[:T, 2 * s + 1] = y_CP[:, s].imag
                eout[dl[2 * s]:dl[2 * s] + T, 2 * s] = x_CP[:, s].real
                eout[dl[2 * s + 1]:dl[2 * s + 1] + T, 2 * s + 1] = x_CP[:, s].imag
            return ein, eout

        scores = np.zeros(table.shape[0])
        ref = x_CP[IsiDuration - 1:, :]
        io = [build(row) for row in range(table.shape[0])]
        if hasattr(esn, "fit_predict_many"):
            # all table rows in one batched harvest / solve / predict (same noise order, same results)
            outs = esn.fit_predict_many([(ein, eout, lo[row] + CyclicPrefixLen) for row, (ein, eout) in enumerate(io)])
        else:
            outs = []
            for row, (ein, eout) in enumerate(io):
                esn.fit(ein, eout, lo[row] + CyclicPrefixLen)
                outs.append(esn.predict(ein, lo[row] + CyclicPrefixLen, continuation=False))
        for row in range(table.shape[0]):
            dl = table[row]
            out = outs[row]
            for s in range(2):
                a, b = dl[2 * s] - lo[row], dl[2 * s + 1] - lo[row]
                est = out[a:a + N + 1, 2 * s] + 1j * out[b:b + N + 1, 2 * s + 1]
                scores[row] += (np.linalg.norm(est - ref[:, s]) ** 2
                                / np.linalg.norm(ref[:, s]) ** 2)
        Delay_Idx = np.argmin(scores)
        Delay_Idx = 3                                  # the reference overrides the argmin (:159)
        print(scores)
        NMSE_ESN = np.amin(scores)
        Delay = table[Delay_Idx, :]
        ESN_input, ESN_output = build(Delay_Idx)
        nForgetPoints = lo[Delay_Idx] + CyclicPrefixLen
        esn.fit(ESN_input, ESN_output, nForgetPoints)
        return [ESN_input, ESN_output, esn, Delay, Delay_Idx, lo[Delay_Idx], hi[Delay_Idx],
                nForgetPoints, NMSE_ESN]
