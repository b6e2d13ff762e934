"""Pin the CPU oracle (oracle/esn_oracle.py) against golden vectors produced by
the live reference (tests/golden/make_golden.py).  CPU only."""
import math

import numpy as np
import pytest

import cases
from conftest import rel_err
from oracle import esn_oracle as orc


@pytest.mark.parametrize("name", list(cases.ESN_CASES))
def test_fit_predict_matches_reference(golden, name):
    c = cases.ESN_CASES[name]
    kw = cases.esn_kwargs(c)
    esn = orc.OracleESN(**kw)
    u, y = cases.esn_io(c, 0)
    u2, _ = cases.esn_io(c, 1)
    pred_train = esn.fit(u, y, c["transient"])
    chk = np.array([esn.W.sum(), np.abs(esn.W).sum(), esn.W_in.sum(), esn.W_feedb.sum()])
    assert np.array_equal(chk, golden[name + "/W_checksum"])          # same weights, bit for bit
    # SURVEY §4 identity 2: restated harvest + pinv reproduces W_out (bit-for-bit on this
    # BLAS; allow round-off so the test survives another BLAS build)
    assert rel_err(esn.W_out, golden[name + "/W_out"]) < 1e-9
    assert rel_err(pred_train, golden[name + "/pred_train"]) < 1e-9
    assert rel_err(esn.laststate, golden[name + "/laststate"]) < 1e-12
    assert np.array_equal(esn.lastinput, golden[name + "/lastinput"])
    assert rel_err(esn.lastoutput, golden[name + "/lastoutput"]) < 1e-15
    pred = esn.predict(u2, c["transient"], continuation=c["continuation"])
    assert pred.shape == golden[name + "/predict"].shape
    # SURVEY §4 identity 1: predict consumes one rand(N_res) row per step
    assert rel_err(pred, golden[name + "/predict"]) < 1e-8
    # identity 3: predict leaves laststate alone
    assert np.array_equal(esn.laststate, golden[name + "/laststate_after_predict"]) or \
        rel_err(esn.laststate, golden[name + "/laststate_after_predict"]) < 1e-12
    assert int(golden[name + "/fit_draws_ok"][0]) == 1


@pytest.mark.parametrize("name", list(cases.TRAINER_CASES))
@pytest.mark.parametrize("flag", [0, 1])
def test_generic_trainer_matches_reference(golden, name, flag):
    c = cases.TRAINER_CASES[name]
    blk = orc.synth_block(c["seed"], c["N"], c["N_t"], c["N_r"], c["m"], c["ebno"], 2,
                          isi_duration=c["isi"])
    cp = c["isi"] - 1
    maxd = int(math.ceil(c["isi"] / 2) + 2)
    kw = cases.trainer_esn_kwargs(c, blk["var_x"])
    kw.pop("feedback_scaling")
    esn = orc.OracleESN(**kw)
    res = orc.train_generic(esn, flag, 0, maxd, cp, c["N"], c["N_t"], c["N_r"], c["isi"],
                            blk["pilot"]["y_CP"], blk["pilot"]["x_CP"])
    key = f"{name}/flag{flag}"
    assert np.array_equal(res[0], golden[key + "/ESN_input"])
    assert np.array_equal(res[1], golden[key + "/ESN_output"])
    assert np.array_equal(res[3], golden[key + "/Delay"])
    assert [res[4], res[5], res[6], res[7]] == list(golden[key + "/scalars"])
    assert rel_err(res[2].W_out, golden[key + "/W_out"]) < 1e-7
    assert abs(res[8] - golden[key + "/NMSE"][0]) < 1e-6 * abs(golden[key + "/NMSE"][0])
    ein = orc.pack_rx(blk["data"][0]["y_CP"], int(res[6]))
    det = res[2].predict(ein, res[7], continuation=False)
    assert rel_err(det, golden[key + "/detect"]) < 1e-6


def test_legacy_trainer_matches_reference(golden):
    name = "gen_2x2"
    c = cases.TRAINER_CASES[name]
    blk = orc.synth_block(c["seed"], c["N"], c["N_t"], c["N_r"], c["m"], c["ebno"], 2,
                          isi_duration=c["isi"])
    cp = c["isi"] - 1
    maxd = int(math.ceil(c["isi"] / 2) + 2)
    kw = cases.trainer_esn_kwargs(c, blk["var_x"])
    kw.pop("feedback_scaling")
    esn = orc.OracleESN(**kw)
    res = orc.train_legacy_2x2(esn, 0, 0, maxd, cp, c["N"], c["N_t"], c["N_r"], c["isi"],
                               blk["pilot"]["y_CP"], blk["pilot"]["x_CP"])
    key = f"{name}/legacy"
    assert np.array_equal(res[0], golden[key + "/ESN_input"])
    assert np.array_equal(res[1], golden[key + "/ESN_output"])
    assert np.array_equal(res[3], golden[key + "/Delay"])
    assert [int(res[4]), int(res[5]), int(res[6]), int(res[7])] == list(golden[key + "/scalars"])
    assert rel_err(res[2].W_out, golden[key + "/W_out"]) < 1e-7
    assert abs(res[8] - golden[key + "/NMSE"][0]) < 1e-6 * abs(golden[key + "/NMSE"][0])
    with pytest.raises(TypeError):
        orc.train_legacy_2x2(esn, 1, 0, maxd, cp, c["N"], c["N_t"], c["N_r"], c["isi"],
                             blk["pilot"]["y_CP"], blk["pilot"]["x_CP"])


@pytest.mark.parametrize("Bi", [2, 4, 6])
def test_qam_table(golden, Bi):
    const = orc.unit_qam_constellation(Bi)
    assert np.allclose(const, golden[f"qam/{Bi}"], rtol=0, atol=1e-15)
    assert np.allclose(const, golden[f"qam_script/{Bi}"], rtol=0, atol=1e-15)


@pytest.mark.parametrize("name", list(cases.TRAINER_CASES))
def test_baseline_chain_matches_reference(golden, name):
    c = cases.TRAINER_CASES[name]
    N, N_t, N_r, m, isi = c["N"], c["N_t"], c["N_r"], c["m"], c["isi"]
    blk = orc.synth_block(c["seed"] + 100, N, N_t, N_r, m, c["ebno"], 2, isi_duration=isi)
    cp, Pi, No = isi - 1, blk["Pi"], blk["No"]
    Y_LS = orc.rx_fft(blk["pilot"]["y_LS_CP"], cp, N)
    H_LS, H_MMSE = orc.channel_estimate(Y_LS, blk["pilot"]["X_LS"], Pi, No, N, N_t, N_r,
                                        blk["isi_magnitude"], isi)
    assert rel_err(H_LS, golden[f"chain/{name}/H_LS"]) < 1e-12
    assert rel_err(H_MMSE, golden[f"chain/{name}/H_MMSE"]) < 1e-12
    Y = orc.rx_fft(blk["data"][0]["y_CP"], cp, N)
    assert rel_err(Y, golden[f"chain/{name}/Y"]) < 1e-14
    Xzf = orc.equalize(Y, blk["H_true"], math.sqrt(Pi), 1e-12)
    Xls = orc.equalize(Y, H_LS, math.sqrt(Pi), 1e-12)
    Xmm = orc.equalize(Y, H_MMSE, math.sqrt(Pi), No / Pi)
    assert rel_err(Xzf, golden[f"chain/{name}/X_perfzf"]) < 1e-10
    assert rel_err(Xls, golden[f"chain/{name}/X_lszf"]) < 1e-10
    assert rel_err(Xmm, golden[f"chain/{name}/X_mmse"]) < 1e-10
    const = orc.unit_qam_constellation(m)
    for key, X in (("mmse", Xmm), ("perfzf", Xzf)):
        idx = orc.hard_demap_indices(X, const)
        assert np.array_equal(orc.indices_to_bits(idx, m), golden[f"chain/{name}/bits_{key}"])
        # the closed-form slicer agrees with the nearest-point search
        assert np.array_equal(orc.slicer_indices(X, m), idx)
    Xe = orc.esn_output_to_freq(golden[f"chain/{name}/esn_out_time"], N, N_t, Pi)
    assert rel_err(Xe, golden[f"chain/{name}/esn_out_freq"]) < 1e-14
    idx = orc.hard_demap_indices(Xe, const)
    assert np.array_equal(orc.indices_to_bits(idx, m), golden[f"chain/{name}/bits_esn"])
    assert np.array_equal(orc.slicer_indices(Xe, m), idx)


def test_bits_roundtrip():
    rng = np.random.RandomState(3)
    for m in (2, 4, 6):
        bits = (rng.rand(16 * m, 3) > 0.5).astype(np.int32)
        idx = orc.bits_to_indices(bits, m)
        assert np.array_equal(orc.indices_to_bits(idx, m), bits)
        const = orc.unit_qam_constellation(m)
        assert np.array_equal(orc.slicer_indices(const[idx], m), idx)
        assert np.array_equal(orc.hard_demap_indices(const[idx], const), idx)


# --------------------------------------------------------------------------
# soft outputs: sigma2 estimate, max-log LLRs, logistic calibration
# --------------------------------------------------------------------------
@pytest.fixture(scope="module")
def soft_golden():
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "soft_golden.npz"))


@pytest.mark.parametrize("m", [2, 4, 6])
def test_oracle_soft_outputs_match_reference(soft_golden, m):
    g = soft_golden
    X, idx = g[f"soft/{m}/X_hat"], g[f"soft/{m}/tx_idx"]
    const = orc.unit_qam_constellation(m)
    X2, idx2 = cases.soft_frames(m, *cases.SOFT_CASES[m], const)
    assert np.array_equal(idx, idx2) and np.allclose(X, X2, rtol=0, atol=1e-15)
    llr = np.stack([orc.frame_llrs(X[f], m)[0] for f in range(X.shape[0])])
    s2 = np.array([orc.frame_llrs(X[f], m)[1] for f in range(X.shape[0])])
    assert np.allclose(s2, g[f"soft/{m}/sigma2"], rtol=1e-13, atol=0)
    assert np.allclose(llr, g[f"soft/{m}/llr"], rtol=1e-12, atol=1e-12)
    ab = np.array([orc.fit_logreg_1d(llr[:, :, b, :].reshape(-1), ((idx >> b) & 1).reshape(-1).astype(float),
                                     maxiter=400, lr=0.1, l2=1e-3) for b in range(m)])
    assert np.allclose(ab, g[f"soft/{m}/ab"], rtol=1e-12, atol=1e-14)
    # the labels are the LSB-first binary of the point index, and hard decisions = sign of the LLR
    hard = (llr < 0).astype(int)
    idx_hat = orc.hard_demap_indices(X[0], const)
    assert np.array_equal(hard[0], np.moveaxis(((idx_hat[:, :, None] >> np.arange(m)) & 1), 2, 1))
    cal = orc.calibrate_llrs(llr, ab[:, 0], ab[:, 1], 20.0)
    assert cal.shape == llr.shape and np.all(np.abs(cal) <= 20.0)


def test_siso_demo_loop_oracle_matches_reference_counts():
    """BASELINE.json configs[0]: the SISO QPSK / AWGN demo loop (examples/siso_qpsk_awgn.py) driven by the
    oracle ESN reproduces the error counts of the same loop driven by the live reference pyESN
    (tests/golden/make_golden_siso.py), detector by detector and symbol by symbol."""
    import importlib.util
    import os
    from conftest import ROOT
    spec = importlib.util.spec_from_file_location("siso_qpsk_awgn", os.path.join(ROOT, "examples", "siso_qpsk_awgn.py"))
    demo = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(demo)
    g = np.load(os.path.join(ROOT, "tests", "golden", "siso_demo_golden.npz"))
    state = np.random.get_state()
    try:
        r = demo.run(orc.OracleESN, [float(e) for e in g["ebno"]], int(g["symbols"]), int(g["nres"]),
                     seed=int(g["seed"]), keep_first=True)
    finally:
        np.random.set_state(state)
    assert r["bits"] == g["bits"].tolist()
    for k in ("ESN", "MMSE", "ZF", "LS"):
        assert r[k] == g["err_" + k].tolist(), k
    assert np.array_equal(np.array(r["esn_per_symbol"]), g["esn_per_symbol"])
    assert rel_err(np.array(r["first_xhat"]), g["first_xhat"]) < 1e-9
