"""Result writers in the reference's artefact formats, so that the notebooks
and plots that consume the reference's `results/` files read this engine's output unchanged.
Formats follow system_model_2/Demo_MIMO_4x8_Sionna_CDL_ESN_v2.py:
  results_ber.csv  header EbNo(dB),ESN_uncoded,MMSE_uncoded,ESN_coded,MMSE_coded, one row per SNR point,
                   the SNR as an int, BERs as Python floats (:636-641);
  results_ber.pkl  dict EBN0 / BER_uncoded{ESN,MMSE} / BER_coded{ESN_calLLR,MMSE_calLLR} [/ meta] (:554-589, :620-633);
  LLR_calibration_params_EbNo<snr>dB.txt  'bit, a_esn, b_esn, a_mmse, b_mmse' and one row per bit with
                   four decimals (:532-535).
Host-side file output only; nothing here touches the GPU."""
import csv
import os
import pickle

import numpy as np


def _col(v, n):
    """A BER column as a list of n Python floats (None -> zeros, like the reference's np.zeros arrays)."""
    if v is None:
        return [0.0] * n
    v = [float(x) for x in np.asarray(v, dtype=np.float64).reshape(-1)]
    if len(v) != n:
        raise ValueError(f"expected {n} SNR points, got {len(v)}")
    return v


def results_bundle(ebno_db, esn_uncoded, mmse_uncoded, esn_coded=None, mmse_coded=None, meta=None):
    """The dict the reference pickles (keys EBN0, BER_uncoded, BER_coded, optional meta)."""
    ebno = [x.item() if hasattr(x, "item") else x for x in np.asarray(ebno_db).reshape(-1)]
    n = len(ebno)
    out = {
        "EBN0": ebno,
        "BER_uncoded": {"ESN": _col(esn_uncoded, n), "MMSE": _col(mmse_uncoded, n)},
        "BER_coded": {"ESN_calLLR": _col(esn_coded, n), "MMSE_calLLR": _col(mmse_coded, n)},
    }
    if meta is not None:
        out["meta"] = meta
    return out


def write_results_csv(path, ebno_db, esn_uncoded, mmse_uncoded, esn_coded=None, mmse_coded=None):
    """results_ber.csv exactly as the reference writes it."""
    b = results_bundle(ebno_db, esn_uncoded, mmse_uncoded, esn_coded, mmse_coded)
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["EbNo(dB)", "ESN_uncoded", "MMSE_uncoded", "ESN_coded", "MMSE_coded"])
        for i, snr in enumerate(b["EBN0"]):
            w.writerow([int(snr), b["BER_uncoded"]["ESN"][i], b["BER_uncoded"]["MMSE"][i],
                        b["BER_coded"]["ESN_calLLR"][i], b["BER_coded"]["MMSE_calLLR"][i]])
    return path


def read_results_csv(path):
    """Inverse of write_results_csv (also reads the reference's own results_ber.csv)."""
    with open(path, newline="") as f:
        rows = list(csv.reader(f))
    if rows[0] != ["EbNo(dB)", "ESN_uncoded", "MMSE_uncoded", "ESN_coded", "MMSE_coded"]:
        raise ValueError("not a results_ber.csv")
    cols = list(zip(*rows[1:])) if len(rows) > 1 else [[], [], [], [], []]
    return results_bundle([int(x) for x in cols[0]], *[[float(x) for x in c] for c in cols[1:]])


def write_results_pkl(path, bundle):
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "wb") as f:
        pickle.dump(bundle, f)
    return path


def write_llr_calibration(path, ab_esn, ab_mmse):
    """LLR_calibration_params_EbNo<snr>dB.txt; ab_* = [m, 2] (a, b) per bit position, e.g. from
    esn_b200.ofdm.llr_calibrate."""
    ab_esn = np.asarray(ab_esn, dtype=np.float64).reshape(-1, 2)
    ab_mmse = np.asarray(ab_mmse, dtype=np.float64).reshape(-1, 2)
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "w") as f:
        f.write("bit, a_esn, b_esn, a_mmse, b_mmse\n")
        for b in range(ab_esn.shape[0]):
            f.write(f"{b}, {ab_esn[b, 0]:.4f}, {ab_esn[b, 1]:.4f}, {ab_mmse[b, 0]:.4f}, {ab_mmse[b, 1]:.4f}\n")
    return path
