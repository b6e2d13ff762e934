"""GPU parity tests of the resident tensor-core kernel with the CUDA-core readout (csrc/recurrence_tcr.cu)
against the CPU oracle: the ladder of the other tensor-core kernels, cfg3 with a readout per 18 frames, harvest,
continuation, both output widths.  Tolerances as in BASELINE.json: states 1e-5 relative, outputs 1e-4 relative."""
import numpy as np
import pytest
import torch

import cases
from conftest import rel_err
from oracle import esn_oracle as orc
from test_gpu_tc import _check, _cuda, _setup

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import __graft_entry__ as g
    g.build()


@pytest.mark.parametrize("n_res,T", [(128, 1), (128, 2), (128, 5), (300, 7), (512, 3)])
def test_tcr_ladder_small_steps(n_res, T):
    """One 256-neuron group (N = 128) and two (300, 512), no noise: step 1 = input block + UMMAs + epilogue +
    CUDA-core readout, step 2 = the in-place state rewrite, 5 steps = feedback."""
    rng, Ws, aff, eng = _setup(n_res, 4, 4, seed=3, noise=0.0)
    B = 70
    us = rng.randn(B, T, 4)
    W_outs = rng.randn(1, 4, n_res + 4) * 1e-6
    _check(eng, Ws, aff, us, W_outs, None, T, 0, 0.0, None, precision="tcr")


def test_tcr_cfg3_a_readout_per_cta():
    """4x8 / 512 neurons / T = 522: a new readout every 64 frames (both CTAs of a pair with their own), ragged last
    tile, host-supplied state noise, oracle-trained readouts."""
    c = cases.ESN_CASES["cfg3_4x8_n512"]
    rng, Ws, aff, eng = _setup(512, 16, 8, seed=42, noise=0.001, in_scale=0.005)
    B, T, G = 128 + 70, c["T"], 11
    us = rng.randn(B, T, 16)
    W_outs = []
    for g in range(G):
        if g < 3:
            u, y = cases.esn_io(c, 20 + g)
            W_outs.append(orc.fit(Ws[0], Ws[1], Ws[2], u, y, 10, 0.001, rng.rand(T - 1, 512), **aff)["W_out"])
        else:
            W_outs.append(W_outs[g % 3] * (1.0 + 0.1 * g))
    W_outs = np.stack(W_outs)
    gid = (np.arange(B) // 64) % G
    uni = rng.rand(B, T, 512)
    ws, wy = _check(eng, Ws, aff, us, W_outs, gid, T, 10, 0.001, uni, precision="tcr",
                    frames=[0, 17, 31, 32, 62, 63, 64, 71, 72, 127, 128, 143, 191, 192, 197])
    print("tcr cfg3: worst state err %.2e, output err %.2e" % (ws, wy))
    bad = gid.copy()
    bad[70] = 5
    with pytest.raises(Exception):
        eng.predict(_cuda(us), _cuda(W_outs), transient=10, group_ids=_cuda(bad.astype(np.int32)), precision="tcr")
    assert ws < 4e-6                     # split accumulators + truncation-bias gain (the first resident kernel: 9e-6)


@pytest.mark.parametrize("n_res", [100, 512])
def test_tcr_harvest_matches_oracle(n_res):
    rng, Ws, aff, eng = _setup(n_res, 16, 8, seed=5, noise=0.001, in_scale=0.005)
    B, T, N = 70, 90, n_res
    us, ts = rng.randn(B, T, 16), rng.randn(B, T, 8)
    uni = rng.rand(B, T - 1, N)
    ext = eng.harvest(_cuda(us), _cuda(ts), precision="tcr", noise_uniforms=_cuda(uni)).double().cpu().numpy()
    for b in (0, 63, 64, 69):
        r = orc.fit(Ws[0], Ws[1], Ws[2], us[b], ts[b], 1, 0.001, uni[b], **aff)
        assert rel_err(ext[b, :, :N], r["states"]) < 1e-5
        assert rel_err(ext[b, :, N:], r["in_s"]) < 1e-6
        assert np.all(ext[b, 0, :N] == 0)


@pytest.mark.parametrize("n_res,n_in,n_out", [(200, 16, 8), (300, 5, 3), (512, 16, 8)])
def test_tcr_widest_io_continuation_device_noise(n_res, n_in, n_out):
    """The widest shape the kernel takes (16 inputs, 8 outputs), an odd one, one and two neuron groups: explicit x0 / y0,
    a feedback strong enough to matter, device noise."""
    from esn_b200.noise import device_noise_uniforms
    rng, Ws, aff, eng = _setup(n_res, n_in, n_out, seed=11, noise=0.001, in_scale=0.02, t_scale=2e-2)
    assert eng.tcr_supported()
    B, T, N = 139, 24, n_res
    us = rng.randn(B, T, n_in)
    W_outs = rng.randn(3, n_out, N + n_in) * 2e-3
    gid = (np.arange(B) // 64) % 3
    x0, y0 = rng.randn(B, N) * 0.1, rng.randn(B, n_out) * 0.02
    seed = 31
    uni = device_noise_uniforms(seed, B, T, N)
    y, ext = eng.predict_tcr(_cuda(us), _cuda(W_outs), transient=1, group_ids=gid, x0=_cuda(x0), y0=_cuda(y0),
                             seed=seed, return_ext=True, y_absmax=0.2)
    y, ext = y.double().cpu().numpy(), ext.double().cpu().numpy()
    fb = 0.0
    for b in (0, 31, 64, 74, 127, 128, 138):
        ref, st = orc.predict(Ws[0], Ws[1], Ws[2], W_outs[gid[b]], us[b], 1, 0.001, uni[b], x0=x0[b], y0=y0[b],
                              return_states=True, **aff)
        assert rel_err(ext[b, :, :N], st) < 1e-5
        assert rel_err(y[b], ref) < 1e-4
        nofb = orc.predict(Ws[0], Ws[1], Ws[2], W_outs[gid[b]], us[b], 1, 0.001, uni[b], x0=x0[b], y0=None,
                           **{**aff, "teacher_forcing": False})
        fb = max(fb, rel_err(nofb, ref))
    assert fb > 1e-2
    # wider I/O stays on the streamed-state kernel
    _, _, _, big = _setup(400, 24, 16, seed=1)
    assert not big.tcr_supported() and big.tcs_supported()


def test_tcr_device_noise_no_feedback_and_small_io():
    """teacher_forcing = False with the device counter noise (restated on the host for the oracle), and the padded
    reservoirs / small I/O shapes of the 2x2 and SISO demos."""
    from esn_b200.noise import device_noise_uniforms
    rng, Ws, aff, eng = _setup(256, 16, 8, seed=5, noise=0.001, feedback=False)
    B, T = 64, 40
    us = rng.randn(B, T, 16)
    W_outs = rng.randn(1, 8, 272) * 1e-5
    seed = 77
    y = eng.predict(_cuda(us), _cuda(W_outs), transient=3, precision="tcr", seed=seed).double().cpu().numpy()
    uni = device_noise_uniforms(seed, B, T, 256)
    for b in (0, 31, 63):
        ref = orc.predict(Ws[0], Ws[1], Ws[2], W_outs[0], us[b], 3, 0.001, uni[b], **aff)
        assert rel_err(y[b], ref) < 1e-4
    for n_res, n_in, n_out in ((100, 4, 4), (200, 2, 2), (300, 4, 4)):
        rng, Ws, aff, eng = _setup(n_res, n_in, n_out, seed=n_res, noise=0.001)
        B, T = 20, 30
        us = rng.randn(B, T, n_in)
        W_outs = rng.randn(1, n_out, n_res + n_in) * 1e-6
        uni = rng.rand(B, T, n_res)
        _check(eng, Ws, aff, us, W_outs, None, T, 2, 0.001, uni, precision="tcr")


def test_tcr_gain_setter_round_trip():
    """esn_tc_set_acc_k0: 0 switches the compensation off (states move by the known bias), a negative value restores
    the calibrated default."""
    from esn_b200._lib import load
    lib = load()
    rng, Ws, aff, eng = _setup(512, 16, 8, seed=42, noise=0.0, in_scale=0.005)
    B, T = 64, 200
    us, ts = rng.randn(B, T, 16), rng.randn(B, T, 8)
    k0 = lib.esn_tc_acc_k0()
    assert 1e-8 < k0 < 2e-8
    on = eng.harvest(_cuda(us), _cuda(ts), precision="tcr").double()
    assert lib.esn_tc_set_acc_k0(0.0) == k0
    off = eng.harvest(_cuda(us), _cuda(ts), precision="tcr").double()
    assert lib.esn_tc_set_acc_k0(-1.0) == 0.0 and lib.esn_tc_acc_k0() == k0
    ref = eng.harvest(_cuda(us).double(), _cuda(ts).double(), precision="fp64")
    e_on = float((on - ref).abs().max() / ref.abs().max())
    e_off = float((off - ref).abs().max() / ref.abs().max())
    print("tcr states vs fp64, gain on / off: %.2e / %.2e" % (e_on, e_off))
    assert e_on < 2e-6 and e_on < 0.6 * e_off
