#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by running the LIVE,
UNMODIFIED reference (`/root/reference/libs`, read-only) on the seeded inputs of
`cases.py`.  Runs only in the build container (the GPU box has no
/root/reference); the resulting `.npz` files are committed.

    python tests/golden/make_golden.py

The chain functions of the demo scripts (`equalize_zf`, `hard_bits_from_syms`,
the inline LS/MMSE channel-estimation block, ...) cannot be imported because
the scripts run a whole simulation at import time and need matplotlib/pyldpc.
They are executed here straight from the reference's source text: function
definitions are pulled out by AST, the inline estimation block by line range
(with a sentinel check), and exec'ed in a scratch namespace.  No reference
source is copied into this repository.
"""
import ast
import os
import sys
import textwrap
import math

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, os.path.join(REF, "libs"))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from pyESN import ESN as RefESN                               # noqa: E402
from helper_mimo_esn_generic import trainMIMOESN_generic      # noqa: E402
from HelpFunc import HelpFunc as RefHelpFunc                  # noqa: E402
import cases                                                  # noqa: E402
from oracle import esn_oracle as orc                          # noqa: E402  (only for the workload generator)


def ref_script_functions(path, names):
    src = open(path).read()
    tree = ast.parse(src)
    from scipy import signal, interpolate
    ns = {"np": np, "math": math, "signal": signal, "interpolate": interpolate}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            code = compile(ast.Module(body=[node], type_ignores=[]), path, "exec")
            exec(code, ns)
    return ns


def ref_chanest_block(path, ns_vars):
    lines = open(path).read().split("\n")
    block = lines[315:334]            # file lines 316..334
    assert "H_LS = np.zeros_like(H_true)" in block[0], "reference layout changed"
    assert "H_MMSE[:, nr, tx] = Hmmse_full" in block[-1], "reference layout changed"
    code = textwrap.dedent("\n".join(block))
    exec(compile(code, path + ":316-334", "exec"), ns_vars)
    return ns_vars["H_LS"], ns_vars["H_MMSE"]


def gen_esn_cases(out):
    for name, c in cases.ESN_CASES.items():
        esn = RefESN(**cases.esn_kwargs(c))
        u, y = cases.esn_io(c, 0)
        u2, _ = cases.esn_io(c, 1)
        st0 = esn.random_state_.get_state()
        pred_train = esn.fit(u, y, c["transient"])
        st1 = esn.random_state_.get_state()
        pred = esn.predict(u2, c["transient"], continuation=c["continuation"])
        out[name + "/W_out"] = esn.W_out
        out[name + "/pred_train"] = pred_train
        out[name + "/laststate"] = esn.laststate
        out[name + "/lastinput"] = esn.lastinput
        out[name + "/lastoutput"] = esn.lastoutput
        out[name + "/predict"] = pred
        out[name + "/W_checksum"] = np.array([esn.W.sum(), np.abs(esn.W).sum(),
                                              esn.W_in.sum(), esn.W_feedb.sum()])
        # a second predict on the same object must not change laststate (SURVEY §4 id. 3)
        out[name + "/laststate_after_predict"] = esn.laststate
        # RNG draw accounting: number of doubles consumed by fit and by predict
        def consumed(sa, sb_rng):
            # replay from state sa until we reach the generator's current state
            r = np.random.RandomState()
            r.set_state(sa)
            return r
        r = consumed(st0, None)
        r.rand(c["T"] - 1, c["n_res"])
        out[name + "/fit_draws_ok"] = np.array(
            [int(all(np.array_equal(a, b) if isinstance(a, np.ndarray) else a == b
                     for a, b in zip(r.get_state(), st1)))])
        print("esn case", name, "W_out", esn.W_out.shape, "draws_ok", out[name + "/fit_draws_ok"])


def gen_trainer_cases(out):
    for name, c in cases.TRAINER_CASES.items():
        blk = orc.synth_block(c["seed"], c["N"], c["N_t"], c["N_r"], c["m"], c["ebno"], 2,
                              isi_duration=c["isi"])
        cp = c["isi"] - 1
        maxd = int(math.ceil(c["isi"] / 2) + 2)
        y_CP, x_CP = blk["pilot"]["y_CP"], blk["pilot"]["x_CP"]
        for flag in (0, 1):
            esn = RefESN(**cases.trainer_esn_kwargs(c, blk["var_x"]))
            res = trainMIMOESN_generic(esn, flag, 0, maxd, cp, c["N"], c["N_t"], c["N_r"],
                                       c["isi"], y_CP, x_CP)
            key = f"{name}/flag{flag}"
            out[key + "/ESN_input"] = res[0]
            out[key + "/ESN_output"] = res[1]
            out[key + "/W_out"] = res[2].W_out
            out[key + "/Delay"] = res[3]
            out[key + "/scalars"] = np.array([res[4], res[5], res[6], res[7]], dtype=np.int64)
            out[key + "/NMSE"] = np.array([res[8]])
            # detect the first data frame with the trained ESN (detect-time call shape)
            d = int(res[6])
            ein = orc.pack_rx(blk["data"][0]["y_CP"], d)
            out[key + "/detect"] = res[2].predict(ein, res[7], continuation=False)
            print("trainer", key, "delay", res[5], "nmse", res[8])
        if c["N_t"] == 2 and c["N_r"] == 2:
            import io, contextlib
            esn = RefESN(**cases.trainer_esn_kwargs(c, blk["var_x"]))
            buf = io.StringIO()
            with contextlib.redirect_stdout(buf):
                res = RefHelpFunc.trainMIMOESN(esn, 0, 0, maxd, cp, c["N"], c["N_t"], c["N_r"],
                                               c["isi"], y_CP, x_CP)
            key = f"{name}/legacy"
            out[key + "/ESN_input"] = res[0]
            out[key + "/ESN_output"] = res[1]
            out[key + "/W_out"] = res[2].W_out
            out[key + "/Delay"] = np.asarray(res[3])
            out[key + "/scalars"] = np.array([res[4], res[5], res[6], res[7]], dtype=np.int64)
            out[key + "/NMSE"] = np.array([res[8]])
            print("trainer", key, "nmse", res[8])


def gen_chain_cases(out):
    script = os.path.join(REF, "system_model_2", "OFDM_MIMO_2-2_NBF_LDPC.py")
    fns = ref_script_functions(script, {
        "unit_qam_constellation", "bits_to_grayvec", "equalize_zf", "equalize_mmse",
        "reconstruct_esn_outputs_generic", "hard_bits_from_syms"})
    for Bi in (2, 4, 6):
        out[f"qam/{Bi}"] = RefHelpFunc.UnitQamConstellation(Bi)
        out[f"qam_script/{Bi}"] = fns["unit_qam_constellation"](Bi)
    from scipy import signal, interpolate
    for name, c in cases.TRAINER_CASES.items():
        N, N_t, N_r, m, isi = c["N"], c["N_t"], c["N_r"], c["m"], c["isi"]
        blk = orc.synth_block(c["seed"] + 100, N, N_t, N_r, m, c["ebno"], 2, isi_duration=isi)
        cp, Pi, No = isi - 1, blk["Pi"], blk["No"]
        Y_LS = (1 / N) * np.fft.fft(blk["pilot"]["y_LS_CP"][cp:, :], axis=0)
        ns = dict(np=np, math=math, interpolate=interpolate, signal=signal,
                  H_true=blk["H_true"].copy(), N=N, N_r=N_r, N_t=N_t,
                  IsiMagnitude=blk["isi_magnitude"], IsiDuration=isi, No=No,
                  Pi=np.array([Pi]), jj=0, X_LS=blk["pilot"]["X_LS"], Y_LS=Y_LS)
        H_LS, H_MMSE = ref_chanest_block(script, ns)
        out[f"chain/{name}/H_LS"] = H_LS
        out[f"chain/{name}/H_MMSE"] = H_MMSE
        fr = blk["data"][0]
        Y = (1 / N) * np.fft.fft(fr["y_CP"][cp:, :], axis=0)
        out[f"chain/{name}/Y"] = Y
        Xzf = np.zeros((N, N_t), dtype=complex)
        Xls = np.zeros((N, N_t), dtype=complex)
        Xmm = np.zeros((N, N_t), dtype=complex)
        for k in range(N):
            Yk = Y[k, :].reshape(N_r, 1)
            Xzf[k] = fns["equalize_zf"](Yk, blk["H_true"][k], math.sqrt(Pi)).reshape(-1)
            Xls[k] = fns["equalize_zf"](Yk, H_LS[k], math.sqrt(Pi)).reshape(-1)
            Xmm[k] = fns["equalize_mmse"](Yk, H_MMSE[k], math.sqrt(Pi), No / Pi).reshape(-1)
        out[f"chain/{name}/X_perfzf"] = Xzf
        out[f"chain/{name}/X_lszf"] = Xls
        out[f"chain/{name}/X_mmse"] = Xmm
        const = fns["unit_qam_constellation"](m)
        p2 = np.power(2, np.arange(m)).reshape((1, -1))
        out[f"chain/{name}/bits_mmse"] = fns["hard_bits_from_syms"](Xmm, const, m, p2)
        out[f"chain/{name}/bits_perfzf"] = fns["hard_bits_from_syms"](Xzf, const, m, p2)
        # unpack + FFT of a synthetic ESN output block
        rng = np.random.RandomState(c["seed"] + 5)
        xh = rng.randn(N, 2 * N_t) * math.sqrt(Pi * N)
        lst = fns["reconstruct_esn_outputs_generic"](xh, np.zeros(2 * N_t, dtype=int), 0, N, N_t)
        Xe = np.zeros((N, N_t), dtype=complex)
        for tx in range(N_t):
            Xe[:, tx] = (1 / N) * np.fft.fft(lst[tx]) / math.sqrt(Pi)
        out[f"chain/{name}/esn_out_time"] = xh
        out[f"chain/{name}/esn_out_freq"] = Xe
        out[f"chain/{name}/bits_esn"] = fns["hard_bits_from_syms"](Xe, const, m, p2)
        print("chain", name, "ok")


def gen_soft_cases(out):
    """Soft-output helpers of the CDL demo (LLRs, noise-variance estimate, logistic calibration)."""
    script = os.path.join(REF, "system_model_2", "Demo_MIMO_4x8_Sionna_CDL_ESN_v2.py")
    fns = ref_script_functions(script, {
        "unit_qam_constellation", "bits_to_grayvec", "qam_bit_labels", "qam_llrs_maxlog",
        "est_sigma2_from_decision", "sigmoid", "fit_logreg_1d"})
    for m, (N, N_t, snr_db, frames) in cases.SOFT_CASES.items():
        const = fns["unit_qam_constellation"](m)
        labels = fns["qam_bit_labels"](2 ** m, m)
        X, idx = cases.soft_frames(m, N, N_t, snr_db, frames, const)
        out[f"soft/{m}/X_hat"] = X
        out[f"soft/{m}/tx_idx"] = idx
        llr_all, s2_all = [], []
        for f in range(frames):
            s2 = np.mean([fns["est_sigma2_from_decision"](X[f, :, tx], const) for tx in range(N_t)])
            ll = np.stack([fns["qam_llrs_maxlog"](X[f, :, tx], const, labels, s2) for tx in range(N_t)], axis=2)
            llr_all.append(ll)
            s2_all.append(s2)
        llr_all = np.stack(llr_all)                       # [frames, N, m, N_t]
        out[f"soft/{m}/llr"] = llr_all
        out[f"soft/{m}/sigma2"] = np.array(s2_all)
        ab = np.zeros((m, 2))
        for b in range(m):
            x = llr_all[:, :, b, :].reshape(-1)
            y = ((idx >> b) & 1).reshape(-1).astype(float)
            ab[b] = fns["fit_logreg_1d"](x, y, maxiter=400, lr=0.1, l2=1e-3)
        out[f"soft/{m}/ab"] = ab
        print("soft", m, "sigma2", s2_all[0], "ab", ab[0])


def gen_writer_cases(out):
    """Text the reference's own writer lines produce (CSV :636-641, calibration txt :532-535) and the
    dict it pickles (:620-633), on synthetic curves."""
    import csv, pickle, tempfile
    script = os.path.join(REF, "system_model_2", "Demo_MIMO_4x8_Sionna_CDL_ESN_v2.py")
    lines = open(script).read().split("\n")
    EbNoDB = np.arange(0, 30 + 1, 3).astype(np.int32)
    rng = np.random.RandomState(31)
    curves = {k: 0.5 * np.exp(-rng.rand() * EbNoDB / 6.0) for k in ("ue", "um", "ce", "cm")}
    tmp = tempfile.mkdtemp()
    ns = dict(os=os, csv=csv, pickle=pickle, np=np, OUT_DIR=tmp, outdir=tmp, EbNoDB=EbNoDB,
              BER_uncoded_ESN=curves["ue"], BER_uncoded_MMSE=curves["um"],
              BER_coded_ESN=curves["ce"], BER_coded_MMSE=curves["cm"],
              m=4, ebno_db=12, a_esn=rng.randn(4), b_esn=rng.randn(4) * 0.01,
              a_mmse=rng.randn(4), b_mmse=rng.randn(4) * 0.01)
    blk = lines[635:641]                                  # file lines 636..641
    assert blk[0].startswith("csv_path = ") and "w.writerow([int(snr)" in blk[-1], "reference layout changed"
    exec(compile("\n".join(blk), script + ":636-641", "exec"), ns)
    blk = lines[531:535]                                  # file lines 532..535
    assert "LLR_calibration_params_EbNo" in blk[0] and "a_mmse[b]" in blk[-1], "reference layout changed"
    exec(compile(textwrap.dedent("\n".join(blk)), script + ":532-535", "exec"), ns)
    blk = lines[619:633]                                  # file lines 620..633: the compact bundle
    assert blk[0].startswith("results_ber = {") and blk[-1].strip() == "pickle.dump(results_ber, f)", "reference layout changed"
    exec(compile("\n".join(blk), script + ":620-633", "exec"), ns)
    out["writers/EbNoDB"] = EbNoDB
    for k, v in curves.items():
        out["writers/" + k] = v
    for k in ("a_esn", "b_esn", "a_mmse", "b_mmse"):
        out["writers/" + k] = ns[k]
    out["writers/csv_bytes"] = np.frombuffer(open(os.path.join(tmp, "results_ber.csv"), "rb").read(), dtype=np.uint8)
    out["writers/txt_bytes"] = np.frombuffer(open(os.path.join(tmp, "LLR_calibration_params_EbNo12dB.txt"), "rb").read(), dtype=np.uint8)
    out["writers/pkl_bytes"] = np.frombuffer(open(os.path.join(tmp, "results_ber.pkl"), "rb").read(), dtype=np.uint8)
    print("writers: csv", len(out["writers/csv_bytes"]), "bytes")


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "core"):
        out = {}
        gen_esn_cases(out)
        gen_trainer_cases(out)
        gen_chain_cases(out)
        path = os.path.join(HERE, "reference_golden.npz")
        np.savez_compressed(path, **out)
        print("wrote", path, os.path.getsize(path) // 1024, "KiB,", len(out), "arrays")
    if which in ("all", "soft"):
        out = {}
        gen_soft_cases(out)
        gen_writer_cases(out)
        path = os.path.join(HERE, "soft_golden.npz")
        np.savez_compressed(path, **out)
        print("wrote", path, os.path.getsize(path) // 1024, "KiB,", len(out), "arrays")


if __name__ == "__main__":
    main()
