"""GPU parity tests of the streamed-state tensor-core kernel (csrc/recurrence_tcs.cu) against the CPU oracle:
per-frame readouts (18-frame coherence blocks that ignore tile boundaries) and reservoirs of 600 / 1024 /
2048 neurons.  Tolerances as in BASELINE.json: states 1e-5 relative, outputs 1e-4 relative."""
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


@pytest.mark.parametrize("T", [1, 2, 5])
def test_tcs_ladder_small_steps(T):
    """N = 128 (one 256-neuron group, one pass), no noise: step 1 = input block + one pass of UMMAs + epilogue +
    CUDA-core readout, step 2 = the state round trip through global memory, 5 steps = feedback."""
    rng, Ws, aff, eng = _setup(128, 4, 4, seed=3, noise=0.0)
    B = 70
    us = rng.randn(B, T, 4)
    W_outs = rng.randn(1, 4, 132) * 1e-6
    _check(eng, Ws, aff, us, W_outs, None, T, 0, 0.0, None, precision="tcs")


def test_tcs_cfg3_per_frame_readouts_at_the_reference_cadence():
    """4x8 / 512 neurons / T = 522: a new readout every 18 frames (the demos' L = 19 cadence,
    OFDM_MIMO_2-2_NBF_LDPC.py:151-153, 270), blocks crossing CTA and pair boundaries, ragged last tile,
    host-supplied state noise, readouts trained by the oracle on pilots."""
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
    gid = (np.arange(B) // 18) % G
    uni = rng.rand(B, T, 512)
    ws, wy = _check(eng, Ws, aff, us, W_outs, gid, T, 10, 0.001, uni, precision="tcs",
                    frames=[0, 17, 18, 35, 36, 63, 64, 71, 72, 127, 128, 143, 144, 197])
    print("tcs cfg3 per-frame readouts: worst state err %.2e, output err %.2e" % (ws, wy))


@pytest.mark.parametrize("n_res,T", [(600, 522), (1024, 522), (2048, 300)])
def test_tcs_large_reservoirs(n_res, T):
    """The reference's 4x8 fast demo uses 600 neurons (Demo_MIMO_4x8_ChannelRank_TrainSNR_LDPC_fast.py:142), the
    sweep of BASELINE.json configs[3] goes to 2048: states within 1e-5 of the oracle, two readouts."""
    rng, Ws, aff, eng = _setup(n_res, 16, 8, seed=7, noise=0.001, in_scale=0.005)
    B = 128 + 5
    us = rng.randn(B, T, 16)
    W_outs = rng.randn(2, 8, n_res + 16) * 1e-6
    gid = (np.arange(B) // 18) % 2
    uni = rng.rand(B, T, n_res)
    ws, wy = _check(eng, Ws, aff, us, W_outs, gid, T, 10, 0.001, uni, precision="tcs", frames=[0, 63, 64, 127, 132])
    print("tcs N=%d: worst state err %.2e, output err %.2e" % (n_res, ws, wy))


def test_tcs_harvest_large_reservoir():
    rng, Ws, aff, eng = _setup(1024, 16, 8, seed=5, noise=0.001, in_scale=0.005)
    B, T, N = 70, 90, 1024
    us, ts = rng.randn(B, T, 16), rng.randn(B, T, 8)
    uni = rng.rand(B, T - 1, N)
    ext = eng.harvest(_cuda(us), _cuda(ts), precision="tc", noise_uniforms=_cuda(uni)).double().cpu().numpy()
    for b in (0, 63, 64, 69):
        r = orc.fit(Ws[0], Ws[1], Ws[2], us[b], ts[b], 1, 0.001, uni[b], **aff)
        assert rel_err(ext[b, :, :N], r["states"]) < 1e-5
        assert rel_err(ext[b, :, N:], r["in_s"]) < 1e-6
        assert np.all(ext[b, 0, :N] == 0)


def test_tcs_widest_io_continuation_and_tuning_knobs():
    """n_in = 24 / n_out = 16 (16 readout accumulators), explicit x0 / y0, a feedback strong enough to matter,
    device noise; the ring knobs change the schedule, not the result."""
    from esn_b200.noise import device_noise_uniforms
    rng, Ws, aff, eng = _setup(300, 24, 16, seed=11, noise=0.001, in_scale=0.02, t_scale=2e-2)
    B, T, N = 75, 24, 300
    us = rng.randn(B, T, 24)
    W_outs = rng.randn(3, 16, N + 24) * 2e-3
    gid = np.arange(B) % 3
    x0, y0 = rng.randn(B, N) * 0.1, rng.randn(B, 16) * 0.02
    seed = 31
    uni = device_noise_uniforms(seed, B, T, N)
    rd = eng.tcs_prepare(_cuda(W_outs))
    outs = []
    for tune in (None, dict(accumulators=2, ring_a=2, ring_b=3), dict(accumulators=4)):
        y, ext = eng.predict_tcs(_cuda(us), rd, transient=1, group_ids=gid, x0=_cuda(x0), y0=_cuda(y0), seed=seed,
                                 return_ext=True, y_absmax=0.2, tune=tune)
        outs.append((y.double().cpu().numpy(), ext.double().cpu().numpy()))
    # ring depths change the schedule only; the accumulator count changes the summation order (not the bar)
    assert np.array_equal(outs[1][0], outs[0][0]) and np.array_equal(outs[1][1], outs[0][1])
    assert rel_err(outs[2][1][:, :, :N], outs[0][1][:, :, :N]) < 1e-5
    y, ext = outs[0]
    fb = 0.0
    for b in (0, 31, 64, 74):
        ref, st = orc.predict(Ws[0], Ws[1], Ws[2], W_outs[gid[b]], us[b], 1, 0.001, uni[b], x0=x0[b], y0=y0[b],
                              return_states=True, **aff)
        assert rel_err(ext[b, :, :N], st) < 1e-5
        assert rel_err(y[b], ref) < 1e-4
        nofb = orc.predict(Ws[0], Ws[1], Ws[2], W_outs[gid[b]], us[b], 1, 0.001, uni[b], x0=x0[b], y0=None,
                           **{**aff, "teacher_forcing": False})
        fb = max(fb, rel_err(nofb, ref))
    assert fb > 1e-2


def test_auto_path_is_never_far_from_the_best_explicit_path():
    """precision='auto' (Reservoir.auto_predict_path) against every explicit path on a batch x reservoir grid:
    it must not lose more than 10 % (+ 0.2 ms of timer slack) to the fastest one."""
    def timed(fn):
        fn()
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        return best
    T = 200
    rows = []
    for n_res, n_in, n_out in ((100, 4, 4), (512, 16, 8), (1024, 16, 8)):
        rng, Ws, aff, eng = _setup(n_res, n_in, n_out, seed=1, noise=0.001, in_scale=0.005)
        Wo = _cuda(rng.randn(1, n_out, n_res + n_in) * 1e-6)
        for B in (8, 64, 256, 1024, 4096):
            u = torch.randn(B, T, n_in, device="cuda")
            t = {"fp32": timed(lambda: eng.predict(u, Wo, transient=10, precision="fp32", seed=1)),
                 "tcs": timed(lambda: eng.predict(u, Wo, transient=10, precision="tcs", seed=1))}
            if eng.tc_supported():
                t["tc"] = timed(lambda: eng.predict(u, Wo, transient=10, precision="tc", seed=1))
            auto = eng.auto_predict_path(B, None)
            rows.append((n_res, B, auto, t))
            assert t[auto] <= 1.1 * min(t.values()) + 0.2, (n_res, B, auto, t)
    print("auto path:", [(n, b, a) for n, b, a, _ in rows])
