"""GPU parity tests (run on the B200 box): the CUDA path, called through the
drop-in API / the C ABI, against the CPU oracle and the reference's golden
vectors on identical seeds.

Tolerances (BASELINE.json north_star): reservoir states 1e-5 relative, readout
weights 1e-4 relative, detected symbol indices bit-exact except symbols within
1e-5 of a decision boundary (counted).  The fp64 path is held to much tighter
bounds where the arithmetic allows.
"""
import math

import numpy as np
import pytest
import torch

import cases
from conftest import rel_err
from oracle import esn_oracle as orc

pytestmark = pytest.mark.gpu

STATE_TOL = 1e-5
WOUT_TOL = 1e-4


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import __graft_entry__ as g
    g.build()


def _cuda(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return t if dtype is None else t.to(dtype)


# --------------------------------------------------------------------------
# drop-in API vs the live reference's golden vectors
# --------------------------------------------------------------------------
@pytest.mark.parametrize("name", list(cases.ESN_CASES))
def test_dropin_fit_predict_matches_reference_golden(golden, name):
    from pyESN import ESN
    c = cases.ESN_CASES[name]
    esn = ESN(**cases.esn_kwargs(c))
    u, y = cases.esn_io(c, 0)
    u2, _ = cases.esn_io(c, 1)
    pred_train = esn.fit(u, y, c["transient"])
    assert pred_train.shape == golden[name + "/pred_train"].shape and pred_train.dtype == np.float64
    assert rel_err(esn.W_out, golden[name + "/W_out"]) < WOUT_TOL
    assert rel_err(pred_train, golden[name + "/pred_train"]) < 1e-5
    assert rel_err(esn.laststate, golden[name + "/laststate"]) < STATE_TOL
    assert np.array_equal(esn.lastinput, golden[name + "/lastinput"])
    assert rel_err(esn.lastoutput, golden[name + "/lastoutput"]) < 1e-14
    before = esn.laststate.copy()
    pred = esn.predict(u2, c["transient"], continuation=c["continuation"])
    assert pred.shape == golden[name + "/predict"].shape
    assert rel_err(pred, golden[name + "/predict"]) < 1e-4
    assert np.array_equal(esn.laststate, before)            # predict has no side effects


@pytest.mark.parametrize("name", list(cases.TRAINER_CASES))
@pytest.mark.parametrize("flag", [0, 1])
def test_generic_trainer_matches_reference_golden(golden, name, flag):
    from pyESN import ESN
    from helper_mimo_esn_generic import trainMIMOESN_generic
    c = cases.TRAINER_CASES[name]
    blk = orc.synth_block(c["seed"], c["N"], c["N_t"], c["N_r"], c["m"], c["ebno"], 2,
                          isi_duration=c["isi"])
    cp, maxd = c["isi"] - 1, int(math.ceil(c["isi"] / 2) + 2)
    esn = ESN(**cases.trainer_esn_kwargs(c, blk["var_x"]))
    res = trainMIMOESN_generic(esn, flag, 0, maxd, cp, c["N"], c["N_t"], c["N_r"], c["isi"],
                               blk["pilot"]["y_CP"], blk["pilot"]["x_CP"])
    key = f"{name}/flag{flag}"
    assert len(res) == 9 and res[2] is esn
    assert np.array_equal(res[0], golden[key + "/ESN_input"])
    assert np.array_equal(res[1], golden[key + "/ESN_output"])
    assert np.array_equal(res[3], golden[key + "/Delay"])
    assert [res[4], res[5], res[6], res[7]] == list(golden[key + "/scalars"])
    assert rel_err(esn.W_out, golden[key + "/W_out"]) < WOUT_TOL
    assert abs(res[8] - golden[key + "/NMSE"][0]) < 1e-4 * abs(golden[key + "/NMSE"][0])
    ein = orc.pack_rx(blk["data"][0]["y_CP"], int(res[6]))
    det = esn.predict(ein, res[7], continuation=False)
    assert rel_err(det, golden[key + "/detect"]) < 1e-4


def test_batched_delay_scan_equals_sequential_calls():
    """ESN.fit_predict_many (one batched harvest / solve / predict for all candidate delays) returns what
    the reference's sequential fit + predict loop returns on the same seed, and leaves the same state."""
    from pyESN import ESN
    from helper_mimo_esn_generic import _pack
    c = cases.TRAINER_CASES["gen_2x2"]
    blk = orc.synth_block(c["seed"], c["N"], c["N_t"], c["N_r"], c["m"], c["ebno"], 2, isi_duration=c["isi"])
    cp = c["isi"] - 1
    y_CP, x_CP = blk["pilot"]["y_CP"], blk["pilot"]["x_CP"]
    cand = [0, 1, 2, 3, 5]
    io = [_pack(y_CP, x_CP, d, c["N"] + cp, c["N_t"], c["N_r"]) for d in cand]
    a = ESN(**cases.trainer_esn_kwargs(c, blk["var_x"]))
    b = ESN(**cases.trainer_esn_kwargs(c, blk["var_x"]))
    seq = []
    for (xi, xo), d in zip(io, cand):
        a.fit(xi, xo, d + cp)
        seq.append(a.predict(xi, d + cp, continuation=False))
    bat = b.fit_predict_many([(xi, xo, d + cp) for (xi, xo), d in zip(io, cand)])
    for s_, b_ in zip(seq, bat):
        assert s_.shape == b_.shape and rel_err(b_, s_) < 1e-9
    assert rel_err(b.W_out, a.W_out) < 1e-9 and rel_err(b.laststate, a.laststate) < 1e-12
    assert np.array_equal(a.random_state_.rand(3), b.random_state_.rand(3))      # generators stayed in step


def test_legacy_trainer_matches_reference_golden(golden, capsys):
    from pyESN import ESN
    from HelpFunc import HelpFunc
    name = "gen_2x2"
    c = cases.TRAINER_CASES[name]
    blk = orc.synth_block(c["seed"], c["N"], c["N_t"], c["N_r"], c["m"], c["ebno"], 2,
                          isi_duration=c["isi"])
    cp, maxd = c["isi"] - 1, int(math.ceil(c["isi"] / 2) + 2)
    esn = ESN(**cases.trainer_esn_kwargs(c, blk["var_x"]))
    res = HelpFunc.trainMIMOESN(esn, 0, 0, maxd, cp, c["N"], c["N_t"], c["N_r"], c["isi"],
                                blk["pilot"]["y_CP"], blk["pilot"]["x_CP"])
    assert "[" in capsys.readouterr().out                    # the NMSE vector is printed, as the reference
    key = f"{name}/legacy"
    assert np.array_equal(res[0], golden[key + "/ESN_input"])
    assert np.array_equal(res[1], golden[key + "/ESN_output"])
    assert np.array_equal(res[3], golden[key + "/Delay"])
    assert [int(res[4]), int(res[5]), int(res[6]), int(res[7])] == list(golden[key + "/scalars"])
    assert rel_err(esn.W_out, golden[key + "/W_out"]) < WOUT_TOL
    assert abs(res[8] - golden[key + "/NMSE"][0]) < 1e-4 * abs(golden[key + "/NMSE"][0])


# --------------------------------------------------------------------------
# kernels vs the oracle: states, readout, batched / grouped predict
# --------------------------------------------------------------------------
def _oracle_esn(c):
    kw = cases.esn_kwargs(c)
    return orc.OracleESN(**kw), kw


@pytest.mark.parametrize("name", ["mimo2x2_small", "mimo4x8_small", "nofeedback_small", "cfg3_4x8_n512"])
@pytest.mark.parametrize("precision", ["fp64", "fp32"])
def test_harvest_states_match_oracle(name, precision):
    from esn_b200 import Reservoir
    c = cases.ESN_CASES[name]
    o, kw = _oracle_esn(c)
    B = 3
    us = np.stack([cases.esn_io(c, i)[0] for i in range(B)])
    ys = np.stack([cases.esn_io(c, i)[1] for i in range(B)])
    uni = np.random.RandomState(77).rand(B, c["T"] - 1, c["n_res"])
    eng = Reservoir(o.W, o.W_in, o.W_feedb, kw["input_scaling"], kw["input_shift"],
                    kw["teacher_scaling"], kw["teacher_shift"], c["noise"], c["teacher_forcing"])
    ext = eng.harvest(_cuda(us), _cuda(ys), precision=precision, noise_uniforms=_cuda(uni)).double().cpu().numpy()
    for b in range(B):
        r = orc.fit(o.W, o.W_in, o.W_feedb, us[b], ys[b], c["transient"], c["noise"], uni[b],
                    input_scaling=kw["input_scaling"], input_shift=kw["input_shift"],
                    teacher_scaling=kw["teacher_scaling"], teacher_shift=kw["teacher_shift"],
                    teacher_forcing=c["teacher_forcing"])
        tol = 1e-11 if precision == "fp64" else STATE_TOL
        assert rel_err(ext[b, :, :c["n_res"]], r["states"]) < tol
        assert rel_err(ext[b, :, c["n_res"]:], r["in_s"]) < (1e-15 if precision == "fp64" else 1e-6)
        assert np.all(ext[b, 0, :c["n_res"]] == 0)


@pytest.mark.parametrize("name", ["mimo2x2_small", "mimo4x8_small", "cfg2_2x2_n100", "cfg3_4x8_n512"])
def test_readout_solve_matches_pinv(name):
    """Gram + Cholesky (primal and dual) vs numpy pinv on the same extended states."""
    from esn_b200 import Reservoir
    c = cases.ESN_CASES[name]
    o, kw = _oracle_esn(c)
    B = 4
    us = np.stack([cases.esn_io(c, i)[0] for i in range(B)])
    ys = np.stack([cases.esn_io(c, i)[1] for i in range(B)])
    uni = np.random.RandomState(5).rand(B, c["T"] - 1, c["n_res"])
    eng = Reservoir(o.W, o.W_in, o.W_feedb, kw["input_scaling"], kw["input_shift"],
                    kw["teacher_scaling"], kw["teacher_shift"], c["noise"], c["teacher_forcing"])
    ext = eng.harvest(_cuda(us), _cuda(ys), precision="fp64", noise_uniforms=_cuda(uni))
    W_out, info = eng.train_readout(ext, _cuda(ys), c["transient"])
    assert info.cpu().tolist() == [0] * B
    E = ext.cpu().numpy()
    for b in range(B):
        teach_s = orc.scale_teacher(ys[b], kw["teacher_scaling"], kw["teacher_shift"])
        ref = (np.linalg.pinv(E[b, c["transient"]:]) @ teach_s[c["transient"]:]).T
        assert rel_err(W_out[b].cpu().numpy(), ref) < 1e-6, (name, b)
    pred = eng.apply_readout(ext, W_out).cpu().numpy()
    for b in range(B):
        ref = orc.unscale_teacher(E[b] @ W_out[b].cpu().numpy().T, kw["teacher_scaling"], kw["teacher_shift"])
        assert rel_err(pred[b], ref) < 1e-12


def test_shared_readout_matches_stacked_pinv():
    """cfg5 semantics: ONE readout over many frames = pinv of the stacked rows."""
    from esn_b200 import Reservoir
    c = cases.ESN_CASES["mimo4x8_small"]
    o, kw = _oracle_esn(c)
    B = 24
    us = np.stack([cases.esn_io(c, i)[0] for i in range(B)])
    ys = np.stack([cases.esn_io(c, i)[1] for i in range(B)])
    eng = Reservoir(o.W, o.W_in, o.W_feedb, kw["input_scaling"], kw["input_shift"],
                    kw["teacher_scaling"], kw["teacher_shift"], c["noise"], c["teacher_forcing"])
    ext = eng.harvest(_cuda(us), _cuda(ys), precision="fp64", seed=9)
    W_out, info = eng.train_readout(ext, _cuda(ys), c["transient"], shared=True)
    assert W_out.shape[0] == 1 and int(info[0]) == 0
    E = ext.cpu().numpy()[:, c["transient"]:].reshape(-1, ext.shape[2])
    D = orc.scale_teacher(ys, kw["teacher_scaling"], kw["teacher_shift"])[:, c["transient"]:].reshape(-1, c["n_out"])
    ref = (np.linalg.pinv(E) @ D).T
    assert rel_err(W_out[0].cpu().numpy(), ref) < 1e-6
    # the two-stage form used across ranks gives the same answer
    G1, R1 = eng.gram(ext[:10], _cuda(ys[:10]), c["transient"])
    G2, R2 = eng.gram(ext[10:], _cuda(ys[10:]), c["transient"])
    W2, info2 = eng.solve_readout(G1 + G2, R1 + R2)
    assert rel_err(W2[0].cpu().numpy(), ref) < 1e-6


def test_rank_deficient_readout_reports_info():
    """All-zero extended states: the kernel-level solve reports LAPACK-style info and a zero pivot ratio;
    the drop-in fit follows the reference, whose pinv of a zero matrix is zero (libs/pyESN.py:191-192
    returns W_out = 0 and raises nothing)."""
    from pyESN import ESN
    esn = ESN(2, 1, n_reservoir=16, random_state=3, noise=0.0)
    u = np.zeros((40, 2))                    # all-zero inputs and teacher -> all-zero states
    eng = esn._engine()
    ext = eng.harvest(u[None], np.zeros((1, 40, 1)), precision="fp64")
    W, info = eng.train_readout(ext, torch.zeros((1, 40, 1), dtype=torch.float64), 0)
    assert int(info[0]) == 1 and not bool(eng.last_pivot_ratio[0] >= eng.PIVOT_RATIO_MIN)
    pred = esn.fit(u, np.zeros((40, 1)))
    ref = orc.OracleESN(2, 1, n_reservoir=16, random_state=3, noise=0.0)
    ref.fit(u, np.zeros((40, 1)))
    assert np.array_equal(esn.W_out, ref.W_out) and np.all(pred == 0)


@pytest.mark.parametrize("name,precision", [("mimo2x2_small", "fp64"), ("mimo2x2_small", "fp32"),
                                            ("mimo4x8_small", "fp32"), ("nofeedback_small", "fp32"),
                                            ("cfg2_2x2_n100", "fp32"), ("cfg3_4x8_n512", "fp32"),
                                            ("cfg3_4x8_n512", "fp64")])
def test_batched_grouped_predict_matches_oracle(name, precision):
    """B frames, G readouts (one per channel realisation), device noise stream:
    states and outputs vs the oracle fed the identical noise."""
    from esn_b200 import Reservoir
    from esn_b200.noise import device_noise_uniforms
    c = cases.ESN_CASES[name]
    o, kw = _oracle_esn(c)
    G, per = 3, 3
    B = G * per - 1                                            # ragged last group
    aff = dict(input_scaling=kw["input_scaling"], input_shift=kw["input_shift"],
               teacher_scaling=kw["teacher_scaling"], teacher_shift=kw["teacher_shift"],
               teacher_forcing=c["teacher_forcing"])
    # train G readouts with the oracle
    rs = np.random.RandomState(21)
    W_outs = []
    for g in range(G):
        u, y = cases.esn_io(c, 10 + g)
        r = orc.fit(o.W, o.W_in, o.W_feedb, u, y, c["transient"], c["noise"],
                    rs.rand(c["T"] - 1, c["n_res"]), **aff)
        W_outs.append(r["W_out"])
    W_outs = np.stack(W_outs)
    us = np.stack([cases.esn_io(c, 50 + i)[0] for i in range(B)])
    gid = np.arange(B) // per
    gid = gid[::-1].copy()                                     # groups not sorted by frame
    seed = 0xABCDEF12345
    uni = device_noise_uniforms(seed, B, c["T"], c["n_res"])
    eng = Reservoir(o.W, o.W_in, o.W_feedb, kw["input_scaling"], kw["input_shift"],
                    kw["teacher_scaling"], kw["teacher_shift"], c["noise"], c["teacher_forcing"])
    y, ext = eng.predict(_cuda(us), _cuda(W_outs), transient=c["transient"],
                         group_ids=_cuda(gid.astype(np.int32)), precision=precision, seed=seed,
                         return_ext=True)
    y, ext = y.double().cpu().numpy(), ext.double().cpu().numpy()
    worst_s = worst_y = 0.0
    for b in range(B):
        ref, st = orc.predict(o.W, o.W_in, o.W_feedb, W_outs[gid[b]], us[b], c["transient"], c["noise"],
                              uni[b], return_states=True, **aff)
        worst_s = max(worst_s, rel_err(ext[b, :, :c["n_res"]], st))
        worst_y = max(worst_y, rel_err(y[b], ref))
    assert worst_s < (1e-10 if precision == "fp64" else STATE_TOL), worst_s
    assert worst_y < (1e-8 if precision == "fp64" else 1e-4), worst_y


def test_predict_is_batch_invariant_and_deterministic():
    """Size-independent properties at the full cfg3 shape: a frame's output does
    not depend on which other frames share its launch or tile, and repeated
    launches are bitwise identical."""
    from esn_b200 import Reservoir
    c = cases.ESN_CASES["cfg3_4x8_n512"]
    o, kw = _oracle_esn(c)
    eng = Reservoir(o.W, o.W_in, o.W_feedb, kw["input_scaling"], kw["input_shift"],
                    kw["teacher_scaling"], kw["teacher_shift"], c["noise"], c["teacher_forcing"])
    rng = np.random.RandomState(4)
    B = 70
    us = _cuda(rng.randn(B, c["T"], c["n_in"]), torch.float32)
    W_out = _cuda(rng.randn(2, c["n_out"], c["n_res"] + c["n_in"]) * 1e-6, torch.float32)
    gid = _cuda((np.arange(B) % 2).astype(np.int32))
    noise = _cuda(rng.rand(B, c["T"], c["n_res"]), torch.float32)
    from esn_b200 import _lib
    lib = _lib.load()
    for limit in (0, 1 << 20):                             # streaming kernel, then the cluster kernel
        old = lib.esn_set_small_batch_limit(limit)
        try:
            y1 = eng.predict(us, W_out, transient=10, group_ids=gid, noise_uniforms=noise)
            y2 = eng.predict(us, W_out, transient=10, group_ids=gid, noise_uniforms=noise)
            assert torch.equal(y1, y2)
            # same kernel, other batch composition / tile
            sub = torch.cat([torch.tensor([5, 17, 69, 33], device="cuda"), torch.arange(20, 60, device="cuda")])
            y3 = eng.predict(us[sub], W_out, transient=10, group_ids=gid[sub], noise_uniforms=noise[sub])
            assert torch.allclose(y3, y1[sub], rtol=1e-5, atol=0)
        finally:
            lib.esn_set_small_batch_limit(old)
        if limit == 0:
            y_stream = y1
    # the two kernels sum in a different order: equal to fp32 accuracy
    assert float((y1 - y_stream).norm() / y_stream.norm()) < 1e-5
    assert y1.shape == (B, c["T"] - 10, c["n_out"]) and torch.isfinite(y1).all()


@pytest.mark.parametrize("n_res,n_in,n_out", [(20, 1, 1), (129, 3, 5), (300, 4, 4), (1024, 16, 8)])
def test_odd_shapes_predict(n_res, n_in, n_out):
    from esn_b200 import Reservoir
    rng = np.random.RandomState(n_res)
    W, W_in, W_fb = orc.init_weights(rng, n_in, n_out, n_res, 0.9, 0.1)
    T, B = 30, 5
    us = rng.randn(B, T, n_in) * 0.1
    W_out = rng.randn(1, n_out, n_res + n_in) * 1e-3
    uni = rng.rand(B, T, n_res)
    eng = Reservoir(W, W_in, W_fb, noise=0.001)
    for precision, tol in (("fp64", 1e-10), ("fp32", 1e-4)):
        y = eng.predict(_cuda(us), _cuda(W_out), precision=precision, noise_uniforms=_cuda(uni)).double().cpu().numpy()
        for b in range(B):
            ref = orc.predict(W, W_in, W_fb, W_out[0], us[b], 0, 0.001, uni[b])
            assert rel_err(y[b], ref) < tol, (precision, b)


@pytest.mark.parametrize("name,kw,T,one_d", [
    ("1-D inputs and outputs (reshape path, libs/pyESN.py:168-171)", dict(n_inputs=1, n_outputs=1, n_reservoir=30), 40, True),
    ("two time steps", dict(n_inputs=2, n_outputs=2, n_reservoir=20), 2, False),
    ("one-row predict", dict(n_inputs=2, n_outputs=2, n_reservoir=20), 12, False),
    ("a single neuron", dict(n_inputs=2, n_outputs=1, n_reservoir=1, sparsity=0.0), 20, False),
    ("513 neurons (past the cluster and tensor-core sizes)", dict(n_inputs=3, n_outputs=2, n_reservoir=513, sparsity=0.2), 30, False),
    ("scalar scalings and shifts (correct_dimensions broadcast)", dict(n_inputs=2, n_outputs=2, n_reservoir=25, input_scaling=0.3,
                                                                       input_shift=0.1, teacher_scaling=2.0, teacher_shift=-0.5), 30, False),
    ("teacher_forcing=False", dict(n_inputs=2, n_outputs=2, n_reservoir=25, teacher_forcing=False), 30, False)])
def test_dropin_edge_shapes_match_oracle(name, kw, T, one_d):
    """Edge shapes of the reference API through the drop-in ESN (fit, then predict with continuation) against
    the oracle on the same seed: degenerate lengths, 1-D arrays, scalar scalings, sizes just past a kernel's
    range."""
    from pyESN import ESN
    a, o = ESN(random_state=5, **kw), orc.OracleESN(random_state=5, **kw)
    rng = np.random.RandomState(1)
    ni, no = kw["n_inputs"], kw["n_outputs"]
    u = rng.randn(T) if one_d else rng.randn(T, ni)
    y = rng.randn(T) if one_d else rng.randn(T, no)
    pa, po = a.fit(u, y), o.fit(u.reshape(T, -1), y.reshape(T, -1))
    assert pa.shape == po.shape == (T, no)
    assert rel_err(pa, po) < 1e-9 and rel_err(a.W_out, o.W_out) < 1e-9
    Tp = 1 if name == "one-row predict" else T
    u2 = rng.randn(Tp) if one_d else rng.randn(Tp, ni)
    qa, qo = a.predict(u2), o.predict(u2.reshape(Tp, -1))
    assert qa.shape == qo.shape == (Tp, no)
    assert rel_err(qa, qo) < 1e-6


@pytest.mark.parametrize("prelude", ["fresh", "odd", "mid-block", "global"])
def test_device_mt19937_continues_the_numpy_stream(prelude):
    """esn_mt19937_uniforms: the doubles of `RandomState.rand` reproduced on the device bit for bit from the
    host generator's state (any position inside a 624-word block, draws spanning many blocks, several draws
    chained), and the state handed back so that the host generator continues as if it had drawn them itself."""
    from esn_b200.noise import DeviceRandomState
    saved = np.random.get_state()
    try:
        if prelude == "global":
            np.random.seed(123)
            rs, ref = np.random.mtrand._rand, np.random.RandomState(123)
        else:
            rs, ref = np.random.RandomState(77), np.random.RandomState(77)
        if prelude == "odd":                               # an odd number of 32-bit outputs consumed
            rs.randint(0, 2 ** 31, size=3, dtype=np.int32), ref.randint(0, 2 ** 31, size=3, dtype=np.int32)
        if prelude == "mid-block":
            rs.rand(1000), ref.rand(1000)
            rs.randn(3), ref.randn(3)                      # leaves a cached gaussian in the state
        d = DeviceRandomState(rs)
        a = d.rand(521, 512)                               # 852 twists
        b = d.rand(7, 3, torch.float32)
        c = d.rand(0, 5)
        e = d.rand(1, 1)
        d.finalize()
        ra, rb, rc, re = ref.rand(521, 512), ref.rand(7, 3), ref.rand(0, 5), ref.rand(1, 1)
        assert np.array_equal(a.cpu().numpy(), ra)
        assert np.array_equal(b.cpu().numpy(), rb.astype(np.float32))
        assert tuple(c.shape) == rc.shape and np.array_equal(e.cpu().numpy(), re)
        assert np.array_equal(rs.rand(2000), ref.rand(2000))            # the host generator continues in step
        assert np.array_equal(rs.randn(5), ref.randn(5))
        sa, sb = rs.get_state(), ref.get_state()
        assert np.array_equal(sa[1], sb[1]) and sa[2:] == sb[2:]
    finally:
        np.random.set_state(saved)


@pytest.mark.parametrize("path", ["cluster", "stream"])
@pytest.mark.parametrize("n_res,n_in,n_out,B", [(40, 2, 2, 1), (100, 4, 4, 3), (200, 2, 2, 5), (300, 16, 8, 9),
                                                  (512, 16, 8, 1), (512, 16, 8, 2), (512, 16, 8, 7)])
def test_small_batch_cluster_and_streaming_kernels_agree_with_oracle(path, n_res, n_in, n_out, B):
    """The two kernels behind esn_recurrence_run (weights resident in a thread-block cluster's shared memory
    for the demos' one-frame calls, weights streamed for large batches) against the oracle on the same
    inputs, in both modes: harvest (libs/pyESN.py:179-182) and predict with continuation state
    (libs/pyESN.py:226-253).  Ragged batches (B not a multiple of the frames-per-cluster tile) included."""
    from esn_b200 import Reservoir
    from esn_b200 import _lib
    lib = _lib.load()
    old = lib.esn_set_small_batch_limit(1 << 20 if path == "cluster" else 0)
    try:
        rng = np.random.RandomState(n_res + B)
        W, W_in, W_fb = orc.init_weights(rng, n_in, n_out, n_res, 0.9, 0.1)
        T = 41
        aff = dict(input_scaling=0.05 * np.ones(n_in), input_shift=0.01 * np.ones(n_in),
                   teacher_scaling=5e-3 * np.ones(n_out), teacher_shift=1e-4 * np.ones(n_out))
        eng = Reservoir(W, W_in, W_fb, aff["input_scaling"], aff["input_shift"], aff["teacher_scaling"],
                        aff["teacher_shift"], 0.001, True)
        us, ys = rng.randn(B, T, n_in), rng.randn(B, T, n_out)
        uni_h, uni_p = rng.rand(B, T - 1, n_res), rng.rand(B, T, n_res)
        W_out = rng.randn(2, n_out, n_res + n_in) * 1e-2
        gid = (np.arange(B) % 2).astype(np.int32)
        x0, y0 = rng.randn(B, n_res) * 0.1, rng.randn(B, n_out) * 5e-3
        for precision, tol in (("fp64", 1e-11), ("fp32", STATE_TOL)):
            ext = eng.harvest(_cuda(us), _cuda(ys), precision=precision, noise_uniforms=_cuda(uni_h))
            y, pext = eng.predict(_cuda(us), _cuda(W_out), transient=3, group_ids=_cuda(gid), precision=precision,
                                  noise_uniforms=_cuda(uni_p), x0=_cuda(x0), y0=_cuda(y0), return_ext=True)
            ext, y, pext = (t.double().cpu().numpy() for t in (ext, y, pext))
            for b in range(B):
                r = orc.fit(W, W_in, W_fb, us[b], ys[b], 0, 0.001, uni_h[b], teacher_forcing=True, **aff)
                assert rel_err(ext[b, :, :n_res], r["states"]) < tol, (precision, b)
                ref, st = orc.predict(W, W_in, W_fb, W_out[gid[b]], us[b], 3, 0.001, uni_p[b], x0=x0[b], y0=y0[b],
                                      return_states=True, **aff)
                assert rel_err(pext[b, :, :n_res], st) < tol, (precision, b)
                assert rel_err(y[b], ref) < (1e-9 if precision == "fp64" else 1e-4), (precision, b)
    finally:
        lib.esn_set_small_batch_limit(old)


# --------------------------------------------------------------------------
# OFDM chain kernels vs the reference's golden vectors
# --------------------------------------------------------------------------
@pytest.mark.parametrize("name", list(cases.TRAINER_CASES))
@pytest.mark.parametrize("dt", ["f64", "f32"])
def test_ofdm_chain_matches_reference_golden(golden, name, dt):
    from esn_b200 import ofdm
    c = cases.TRAINER_CASES[name]
    N, N_t, N_r, m, isi = c["N"], c["N_t"], c["N_r"], c["m"], c["isi"]
    blk = orc.synth_block(c["seed"] + 100, N, N_t, N_r, m, c["ebno"], 2, isi_duration=isi)
    cp, Pi, No = isi - 1, blk["Pi"], blk["No"]
    cd = torch.complex128 if dt == "f64" else torch.complex64
    rd = torch.float64 if dt == "f64" else torch.float32
    tol = 1e-11 if dt == "f64" else 2e-5
    # Rx FFT
    y_cp = _cuda(np.stack([blk["data"][0]["y_CP"], blk["pilot"]["y_LS_CP"]]), cd)
    Y = ofdm.rx_fft(y_cp, N, cp)
    assert rel_err(Y[0].cpu().numpy(), golden[f"chain/{name}/Y"]) < tol
    # channel estimate from the LS pilot
    H_LS, H_MM = ofdm.chanest(Y[1:2], _cuda(blk["pilot"]["X_LS"][None], cd), Pi, blk["isi_magnitude"], isi, No)
    assert rel_err(H_LS[0].cpu().numpy(), golden[f"chain/{name}/H_LS"]) < tol
    assert rel_err(H_MM[0].cpu().numpy(), golden[f"chain/{name}/H_MMSE"]) < tol * 10
    # equalisers (perfect-CSI ZF, LS-ZF, MMSE)
    Ht = _cuda(blk["H_true"][None], cd)
    sp = math.sqrt(Pi)
    X_zf = ofdm.equalize(Y[0:1], Ht, 1e-12, sp)
    X_ls = ofdm.equalize(Y[0:1], H_LS, 1e-12, sp)
    X_mm = ofdm.equalize(Y[0:1], H_MM, No / Pi, sp)
    etol = 1e-8 if dt == "f64" else 5e-3          # ZF with eps=1e-12 is ill-conditioned in fp32
    assert rel_err(X_zf[0].cpu().numpy(), golden[f"chain/{name}/X_perfzf"]) < etol
    assert rel_err(X_ls[0].cpu().numpy(), golden[f"chain/{name}/X_lszf"]) < etol
    assert rel_err(X_mm[0].cpu().numpy(), golden[f"chain/{name}/X_mmse"]) < etol
    # hard decisions + error counting on the reference's own equaliser outputs
    tx_idx = _cuda(blk["data"][0]["idx"][None].astype(np.uint8))
    for key in ("mmse", "perfzf"):
        Xg = golden[f"chain/{name}/X_{key}"]
        idx, counts = ofdm.demap_count(_cuda(Xg[None], cd), m, tx_idx=tx_idx, boundary_eps=1e-5)
        bits = orc.indices_to_bits(idx[0].cpu().numpy().astype(int), m)
        near = orc.boundary_distance(Xg, m) < 1e-5
        ref_bits = golden[f"chain/{name}/bits_{key}"]
        sym_mismatch = (bits != ref_bits).reshape(N, m, N_t).any(axis=1)
        assert not (sym_mismatch & ~near).any()
        if not near.any():
            assert int(counts[0]) == int((ref_bits != blk["data"][0]["bits"]).sum())
    # ESN output -> unpack -> FFT -> demap
    xh = golden[f"chain/{name}/esn_out_time"]
    X, idx, counts = ofdm.unpack_fft_demap(_cuda(xh[None], rd), N, N_t, Pi, m, tx_idx=tx_idx, boundary_eps=1e-5)
    assert rel_err(X[0].cpu().numpy(), golden[f"chain/{name}/esn_out_freq"]) < tol
    bits = orc.indices_to_bits(idx[0].cpu().numpy().astype(int), m)
    near = orc.boundary_distance(golden[f"chain/{name}/esn_out_freq"], m) < 1e-5
    mism = (bits != golden[f"chain/{name}/bits_esn"]).reshape(N, m, N_t).any(axis=1)
    assert not (mism & ~near).any()
    assert int(counts[1]) >= int(near.sum()) - 1


def test_fft_sizes_and_roundtrip_properties():
    """Size-independent properties: Parseval and linearity of the fused FFT at
    every supported size, fp32 and fp64."""
    from esn_b200 import ofdm
    rng = np.random.RandomState(8)
    for N in (2, 64, 512, 4096):
        for rd, cd, tol in ((torch.float64, torch.complex128, 1e-12), (torch.float32, torch.complex64, 3e-6)):
            a = rng.randn(2, N + 3, 2) + 1j * rng.randn(2, N + 3, 2)
            Y = ofdm.rx_fft(_cuda(a, cd), N, 3).cpu().numpy()
            ref = np.fft.fft(a[:, 3:, :], axis=1) / N
            assert rel_err(Y, ref) < tol, (N, rd)


def test_end_to_end_detect_bits_match_oracle():
    """Whole detect path at a small shape: train with the drop-in trainer, detect
    data frames in one batched launch, FFT + demap on the device; symbol
    decisions must equal the oracle's except next to a decision boundary."""
    from pyESN import ESN
    from helper_mimo_esn_generic import trainMIMOESN_generic
    from esn_b200 import ofdm
    c = cases.TRAINER_CASES["gen_2x2"]
    N, N_t, N_r, m, isi = c["N"], c["N_t"], c["N_r"], c["m"], c["isi"]
    n_data = 6
    blk = orc.synth_block(c["seed"] + 7, N, N_t, N_r, m, 30, n_data, isi_duration=isi)
    cp, maxd = isi - 1, int(math.ceil(isi / 2) + 2)
    kw = cases.trainer_esn_kwargs(c, blk["var_x"])
    kw["noise"] = 0.0                                      # decisions then depend on the data only
    gpu, cpu = ESN(**kw), None
    kwo = dict(kw)
    kwo.pop("feedback_scaling")
    cpu = orc.OracleESN(**kwo)
    rg = trainMIMOESN_generic(gpu, 0, 0, maxd, cp, N, N_t, N_r, isi, blk["pilot"]["y_CP"], blk["pilot"]["x_CP"])
    rc = orc.train_generic(cpu, 0, 0, maxd, cp, N, N_t, N_r, isi, blk["pilot"]["y_CP"], blk["pilot"]["x_CP"])
    assert rel_err(gpu.W_out, cpu.W_out) < WOUT_TOL
    d, nforget = int(rg[6]), int(rg[7])
    frames = np.stack([orc.pack_rx(f["y_CP"], d) for f in blk["data"]])
    tx_idx = np.stack([f["idx"] for f in blk["data"]]).astype(np.uint8)
    W_out = torch.from_numpy(gpu.W_out[None]).cuda()
    for precision in ("fp64", "fp32"):
        y = gpu.predict_batched(_cuda(frames), W_out, transient=nforget, precision=precision)
        X, idx, counts = ofdm.unpack_fft_demap(y, N, N_t, blk["Pi"], m, tx_idx=_cuda(tx_idx), boundary_eps=1e-5)
        idx = idx.cpu().numpy().astype(int)
        errs_ref = 0
        for b, f in enumerate(blk["data"]):
            p = cpu.predict(frames[b], nforget, continuation=False)
            Xo = orc.esn_output_to_freq(p, N, N_t, blk["Pi"])
            io = orc.hard_demap_indices(Xo, blk["const"])
            near = orc.boundary_distance(Xo, m) < 1e-5
            assert not ((idx[b] != io) & ~near).any(), (precision, b)
            errs_ref += int((orc.indices_to_bits(io, m) != f["bits"]).sum())
        assert abs(int(counts[0]) - errs_ref) <= 2 * int(counts[1]) + (0 if precision == "fp64" else 2)


# --------------------------------------------------------------------------
# round-1 advisor findings
# --------------------------------------------------------------------------
@pytest.mark.parametrize("n_res,kind", [(500, "steps"), (200, "sin"), (200, "steps")])
def test_fit_noise0_illconditioned_matches_reference_pinv(n_res, kind):
    """noise = 0, one smooth input, long frame: cond(E) = 7e7 (500 neurons, piecewise-constant input: lambda = 0
    normal equations are off by 1e-2 without failing), 3e14 (200 neurons, pure sine: the factorisation breaks
    down) and 5e5 (stays on the Cholesky path), where the reference's SVD pinv (libs/pyESN.py:191-192) returns
    a result every time.  The drop-in fit must follow the reference."""
    from pyESN import ESN
    T = 2000
    kw = dict(n_inputs=1, n_outputs=1, n_reservoir=n_res, spectral_radius=0.95, sparsity=0, noise=0,
              random_state=3)
    rng = np.random.RandomState(8)
    u = np.sin(np.arange(T) / 7.0)[:, None] if kind == "sin" else np.repeat(rng.rand(T // 100), 100)[:, None]
    y = np.roll(u, 2, axis=0) * 0.5 + 0.1 * u ** 2
    gpu, cpu = ESN(**kw), orc.OracleESN(**kw)
    pg, pc = gpu.fit(u, y), cpu.fit(u, y)
    ratio = float(gpu._engine().last_pivot_ratio[0])
    e_w, e_p = rel_err(gpu.W_out, cpu.W_out), rel_err(pg, pc)
    print("noise=0 fit N=%d %s: Cholesky pivot ratio %.2e, W_out rel err vs pinv %.2e, train prediction %.2e"
          % (n_res, kind, ratio, e_w, e_p))
    # cond(E) = 3e14 (sine) is 1 / eps: nothing is well defined beyond "a finite fit of the training rows"
    assert np.isfinite(gpu.W_out).all() and e_p < (1e-2 if kind == "sin" else 1e-6)
    if kind == "steps":                                     # W_out itself is well defined here (cond(E) << 1e15)
        assert e_w < WOUT_TOL
        assert (ratio < gpu._engine().PIVOT_RATIO_MIN) == (n_res == 500)


def test_attributes_changed_after_construction_take_effect():
    """The reference reads noise / scalings / weights live; the cached device engine must follow."""
    from pyESN import ESN
    c = cases.ESN_CASES["mimo2x2_small"]
    kw = cases.esn_kwargs(c)
    u, y = cases.esn_io(c, 0)
    u2, _ = cases.esn_io(c, 1)
    gpu, cpu = ESN(**kw), orc.OracleESN(**kw)
    gpu.fit(u, y, c["transient"]); cpu.fit(u, y, c["transient"])
    for e in (gpu, cpu):
        e.noise = 0
        e.input_scaling = e.input_scaling * 2.0
    pg, pc = gpu.fit(u, y, c["transient"]), cpu.fit(u, y, c["transient"])
    assert rel_err(gpu.W_out, cpu.W_out) < WOUT_TOL and rel_err(pg, pc) < 1e-5
    assert rel_err(gpu.predict(u2, c["transient"], continuation=False),
                   cpu.predict(u2, c["transient"], continuation=False)) < 1e-4
    for e in (gpu, cpu):
        e.W = e.W * 0.5
    gpu.fit(u, y, c["transient"]); cpu.fit(u, y, c["transient"])
    assert rel_err(gpu.W_out, cpu.W_out) < WOUT_TOL


def test_bad_group_ids_are_refused_or_clamped():
    from esn_b200 import Reservoir
    rng = np.random.RandomState(2)
    W, W_in, W_fb = orc.init_weights(rng, 4, 4, 64, 0.9, 0.1)
    eng = Reservoir(W, W_in, W_fb, noise=0.0)
    us = _cuda(rng.randn(6, 20, 4))
    W_outs = _cuda(rng.randn(2, 4, 68) * 1e-3)
    with pytest.raises(ValueError):
        eng.predict(us, W_outs, group_ids=np.array([0, 1, 2, 0, 1, 0]), precision="fp32")
    bad = torch.tensor([0, 1, 7, -3, 1, 0], device="cuda")             # device ids: clamped, never out of bounds
    ok = torch.tensor([0, 1, 1, 0, 1, 0], device="cuda")
    for precision in ("fp32", "fp64"):
        a = eng.predict(us, W_outs, group_ids=bad, precision=precision)
        b = eng.predict(us, W_outs, group_ids=ok, precision=precision)
        assert torch.equal(a, b)


@pytest.mark.parametrize("name", ["mimo4x8_small", "nofeedback_small", "cfg3_4x8_n512"])
def test_large_batch_fp64_harvest_on_the_fp64_tensor_cores_matches_oracle(name):
    """Batches that fill the GPU with 32-frame tiles take `esn_harvest_dmma_kernel` (csrc/recurrence_dmma.cu)
    instead of the streaming SIMT kernel: same E rows as the oracle's teacher-forced loop (libs/pyESN.py:179-182)
    to 1e-11 on sampled frames incl. the ragged last tile, and the same device noise stream as the SIMT kernel."""
    from esn_b200 import Reservoir
    c = cases.ESN_CASES[name]
    o, kw = _oracle_esn(c)
    T = min(c["T"], 48)
    B = 148 * 32 - 7                                        # 148 tiles, the last one ragged
    base_u = np.stack([cases.esn_io(c, i)[0][:T] for i in range(4)])
    base_y = np.stack([cases.esn_io(c, i)[1][:T] for i in range(4)])
    rng = np.random.RandomState(3)
    scale = 1.0 + 0.1 * rng.rand(B, 1, 1)
    us, ys = base_u[np.arange(B) % 4] * scale, base_y[np.arange(B) % 4] * scale
    eng = Reservoir(o.W, o.W_in, o.W_feedb, kw["input_scaling"], kw["input_shift"],
                    kw["teacher_scaling"], kw["teacher_shift"], c["noise"], c["teacher_forcing"])
    pick = [0, 31, 32, 2000, B - 1]
    uni = rng.rand(B, T - 1, c["n_res"])
    ext = eng.harvest(_cuda(us), _cuda(ys), precision="fp64", noise_uniforms=_cuda(uni))
    assert ext.shape == (B, T, c["n_res"] + c["n_in"])
    for b in pick:
        r = orc.fit(o.W, o.W_in, o.W_feedb, us[b], ys[b], 0, c["noise"], uni[b],
                    input_scaling=kw["input_scaling"], input_shift=kw["input_shift"],
                    teacher_scaling=kw["teacher_scaling"], teacher_shift=kw["teacher_shift"],
                    teacher_forcing=c["teacher_forcing"])
        e = ext[b].cpu().numpy()
        assert rel_err(e[:, :c["n_res"]], r["states"]) < 1e-11
        assert rel_err(e[:, c["n_res"]:], r["in_s"]) < 1e-15
        assert np.all(e[0, :c["n_res"]] == 0)
    # device noise stream: a small batch (SIMT / cluster kernel) draws the same numbers for the same frame index
    ext_d = eng.harvest(_cuda(us), _cuda(ys), precision="fp64", seed=11)
    ext_s = eng.harvest(_cuda(us[:40]), _cuda(ys[:40]), precision="fp64", seed=11)
    assert rel_err(ext_d[:40].cpu().numpy(), ext_s.cpu().numpy()) < 1e-11


@pytest.mark.parametrize("n_res,n_in,n_out,per_group", [(512, 16, 8, 0), (512, 16, 8, 64), (100, 4, 4, 0),
                                                         (100, 4, 4, 128), (72, 64, 16, 0), (300, 2, 1, 48)])
def test_large_batch_fp64_predict_on_the_fp64_tensor_cores_matches_oracle(n_res, n_in, n_out, per_group):
    """Free-running fp64 prediction of a GPU-filling batch (16-frame tiles of `esn_harvest_dmma_kernel<.., true>`):
    states and outputs of sampled frames against the oracle's loop (libs/pyESN.py:226-253) with a readout per
    frame group, continuation state and last output, host noise; ragged last tile."""
    from esn_b200 import Reservoir
    rng = np.random.RandomState(n_res)
    W, W_in, W_fb = orc.init_weights(rng, n_in, n_out, n_res, 0.9, 0.1)
    T, B, G = 30, 148 * 16 - 5, 5
    aff = dict(input_scaling=0.05 * np.ones(n_in), input_shift=0.01 * np.ones(n_in),
               teacher_scaling=5e-3 * np.ones(n_out), teacher_shift=1e-4 * np.ones(n_out))
    eng = Reservoir(W, W_in, W_fb, aff["input_scaling"], aff["input_shift"], aff["teacher_scaling"],
                    aff["teacher_shift"], 0.001, True)
    us = rng.randn(B, T, n_in)
    uni = rng.rand(B, T, n_res)
    W_out = rng.randn(G, n_out, n_res + n_in) * 1e-2
    gid = rng.randint(0, G, size=B).astype(np.int32)
    if per_group:                                           # runs of frames per readout: tiles with one readout stage it in shared memory
        gid = ((np.arange(B) // per_group) % G).astype(np.int32)
    x0, y0 = rng.randn(B, n_res) * 0.1, rng.randn(B, n_out) * 5e-3
    y, pext = eng.predict(_cuda(us), _cuda(W_out), transient=3, group_ids=_cuda(gid), precision="fp64",
                          noise_uniforms=_cuda(uni), x0=_cuda(x0), y0=_cuda(y0), return_ext=True)
    assert y.shape == (B, T - 3, n_out)
    for b in (0, 15, 16, 1000, B - 1):
        ref, st = orc.predict(W, W_in, W_fb, W_out[gid[b]], us[b], 3, 0.001, uni[b], x0=x0[b], y0=y0[b],
                              return_states=True, **aff)
        assert rel_err(pext[b, :, :n_res].cpu().numpy(), st) < 1e-11, b
        assert rel_err(y[b].cpu().numpy(), ref) < 1e-9, b


@pytest.mark.parametrize("n_res,B", [(600, 148 * 16 - 3), (1000, 1190), (700, 300)])
def test_wide_reservoirs_fp64_on_the_fp64_tensor_cores_match_oracle(n_res, B):
    """513..1024 neurons in fp64 (the 4x8 fast demo's 600, Demo_MIMO_4x8_ChannelRank_TrainSNR_LDPC_fast.py:142): 16 warps
    of 40 / 48 / 64 neurons cover the reservoir in one pass of `esn_harvest_dmma_kernel`; harvest and free-running
    prediction of sampled frames against the oracle's loops."""
    from esn_b200 import Reservoir
    n_in, n_out, T = 16, 8, 24
    rng = np.random.RandomState(n_res)
    W, W_in, W_fb = orc.init_weights(rng, n_in, n_out, n_res, 0.9, 0.1)
    aff = dict(input_scaling=0.05 * np.ones(n_in), input_shift=0.01 * np.ones(n_in),
               teacher_scaling=5e-3 * np.ones(n_out), teacher_shift=1e-4 * np.ones(n_out))
    eng = Reservoir(W, W_in, W_fb, aff["input_scaling"], aff["input_shift"], aff["teacher_scaling"],
                    aff["teacher_shift"], 0.001, True)
    us, ys = rng.randn(B, T, n_in), rng.randn(B, T, n_out)
    uni_h, uni_p = rng.rand(B, T - 1, n_res), rng.rand(B, T, n_res)
    W_out = rng.randn(3, n_out, n_res + n_in) * 1e-2
    gid = rng.randint(0, 3, size=B).astype(np.int32)
    x0, y0 = rng.randn(B, n_res) * 0.1, rng.randn(B, n_out) * 5e-3
    ext = eng.harvest(_cuda(us), _cuda(ys), precision="fp64", noise_uniforms=_cuda(uni_h))
    y, pext = eng.predict(_cuda(us), _cuda(W_out), transient=2, group_ids=_cuda(gid), precision="fp64",
                          noise_uniforms=_cuda(uni_p), x0=_cuda(x0), y0=_cuda(y0), return_ext=True)
    for b in (0, 7, 8, 17, B - 1):
        r = orc.fit(W, W_in, W_fb, us[b], ys[b], 0, 0.001, uni_h[b], teacher_forcing=True, **aff)
        assert rel_err(ext[b, :, :n_res].cpu().numpy(), r["states"]) < 1e-11, b
        assert rel_err(ext[b, :, n_res:].cpu().numpy(), r["in_s"]) < 1e-15, b
        ref, st = orc.predict(W, W_in, W_fb, W_out[gid[b]], us[b], 2, 0.001, uni_p[b], x0=x0[b], y0=y0[b],
                              return_states=True, **aff)
        assert rel_err(pext[b, :, :n_res].cpu().numpy(), st) < 1e-11, b
        assert rel_err(y[b].cpu().numpy(), ref) < 1e-9, b
