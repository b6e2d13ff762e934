"""Device-resident restatement of the block-fading demo loop (uncoded part) of the reference's System
Model 2 template (system_model_2/OFDM_MIMO_2-2_NBF_LDPC.py:230-474 with USE_LDPC off): per coherence
block one pilot OFDM symbol trains an ESN readout and yields the LS / MMSE channel estimates, the data
symbols of the block are detected by the ESN and by the Perfect-ZF / LS-ZF / MMSE baselines, and bit
errors are counted per detector.

Where the reference walks symbol by symbol in Python, everything here is batched over all blocks of an
SNR point: frame synthesis (`ofdm_synth_frames`), comb-pilot channel estimation (`ofdm_chanest`), readout
training (harvest + Gram + Cholesky), one grouped `predict` over every data frame, equalisation and
slicing + error counting kernels.  One reservoir (fixed `random_state`) serves all blocks, each block
gets its own W_out.  Blocks shard over ranks; counters are summed with one allreduce.
"""
import math

import numpy as np
import torch

from . import dist as D
from . import ofdm

DETECTORS = ("ESN", "Perfect_ZF", "LS_ZF", "MMSE")
# the template's second ESN: trained on the pilot re-sent at a FIXED Eb/N0 (12 dB in the demos), applied to the
# data symbols of the actual Eb/N0 (OFDM_MIMO_2-2_NBF_LDPC.py:182-183, 240-241, 346-367, 440-448)
DETECTORS_TRAIN_FIXED = DETECTORS + ("ESN_trainFixed",)


def isi_profile(isi, device):
    """Exponential power-delay profile of the template (:162-164), normalised to unit sum."""
    mag = torch.exp(-torch.arange(isi, dtype=torch.float64, device=device) / ((isi - 1) / 9))
    return mag / mag.sum()


# 3GPP TR 38.901 Table 7.7.2-2, TDL-B: normalised delays and powers [dB] of the 23 paths (the table the
# reference's CDL demo uses, system_model_2/Demo_MIMO_4x8_Sionna_CDL_ESN_v2.py:127-137)
TDLB_NORM_DELAYS = (0.0000, 0.1072, 0.2155, 0.2095, 0.2870, 0.2986, 0.3752, 0.5055, 0.3681, 0.3697, 0.5700, 0.5283,
                    1.1021, 1.2756, 1.5474, 1.7842, 2.0169, 2.8294, 3.0219, 3.6187, 4.1067, 4.2790, 4.7834)
TDLB_POW_DB = (0.0, -2.2, -4.0, -3.2, -9.8, -1.2, -3.4, -5.2, -7.6, -3.0, -8.9, -9.0, -4.8, -5.7, -7.5, -1.9, -7.6,
               -12.2, -9.8, -11.4, -14.9, -9.2, -11.3)


def tdlb_taps(G, N_r, N_t, isi, fs_hz, ds_ns, generator, device):
    """Block-fading taps [G, N_r, N_t, isi] from the TDL-B profile scaled to an RMS delay spread of `ds_ns`
    (Demo_MIMO_4x8_Sionna_CDL_ESN_v2.py:139-177): every path gets an independent CN(0, p) gain, lands on
    the sample grid with a linear split between the two neighbouring taps, paths beyond `isi` taps are
    dropped, and each link is normalised to unit energy."""
    p = torch.tensor(TDLB_POW_DB, dtype=torch.float64, device=device)
    p = 10 ** (p / 10)
    p = p / p.sum()
    d = torch.tensor(TDLB_NORM_DELAYS, dtype=torch.float64, device=device) * ds_ns * 1e-9 * fs_hz
    i0 = torch.floor(d).long()
    frac = d - i0
    n_path = p.numel()
    g = (torch.randn((G, N_r, N_t, n_path), generator=generator, device=device, dtype=torch.float64)
         + 1j * torch.randn((G, N_r, N_t, n_path), generator=generator, device=device, dtype=torch.float64)) / math.sqrt(2)
    g = g * p.sqrt()
    h = torch.zeros((G, N_r, N_t, isi + 1), dtype=torch.complex128, device=device)
    ok0, ok1 = i0 < isi, (i0 + 1) < isi
    h.index_add_(3, i0[ok0], g[..., ok0] * (1 - frac[ok0]).to(torch.complex128))
    h.index_add_(3, (i0 + 1)[ok1], g[..., ok1] * frac[ok1].to(torch.complex128))
    h = h[..., :isi]
    return h / h.abs().pow(2).sum(dim=3, keepdim=True).sqrt()


def comb_pilot(pil_idx):
    """LS comb pilot: Tx `tx` keeps the pilot symbols on subcarriers tx::N_t, the rest is empty (:287-289)."""
    G, N, N_t = pil_idx.shape
    k = torch.arange(N, device=pil_idx.device)[:, None]
    tx = torch.arange(N_t, device=pil_idx.device)[None, :]
    keep = (k % N_t) == tx
    return torch.where(keep[None], pil_idx, torch.full_like(pil_idx, 255))


def const_table(qam_bits, device, dtype=torch.complex128):
    side = 1 << (qam_bits // 2)
    pam = torch.arange(-(side - 1), side, 2, dtype=torch.float64, device=device)
    c = (pam[:, None] + 1j * pam[None, :]).reshape(-1)
    return (c / math.sqrt(2.0 * (side * side - 1) / 3.0)).to(dtype)


_SIDE = {}


def _side_stream(dev):
    """One side stream per device for the work that runs beside the readout training."""
    key = torch.device(dev).index if torch.device(dev).index is not None else torch.cuda.current_device()
    if key not in _SIDE:
        _SIDE[key] = torch.cuda.Stream(device=key)
    return _SIDE[key]


def channel_frequency_response(taps, N):
    """H_true[g, k, nr, nt] = sum_t taps[g, nr, nt, t] exp(-2 pi i k t / N) (OFDM_MIMO_2-2_NBF_LDPC.py:281-285) through
    the repo's own FFT kernel: the taps of every link are laid out as a zero-padded frame column."""
    G, N_r, N_t, isi = taps.shape
    col = torch.zeros((G, N, N_r * N_t), dtype=torch.complex128, device=taps.device)
    col[:, :isi, :] = taps.to(torch.complex128).permute(0, 3, 1, 2).reshape(G, isi, N_r * N_t)
    H = ofdm.rx_fft(col, N, 0) * N                     # rx_fft carries the 1/N of the receiver side (:428)
    return H.reshape(G, N, N_r, N_t)


def detect_blocks(res, pil_idx, data_idx, block_of_frame, taps, ebno_db, N, qam_bits, isi=8, No=1e-5,
                  clip_db=3.0, delay=None, fit_precision="fp64", detect_precision="fp32", seed=0,
                  noise_pilot=None, noise_data=None, state_noise_seed=1, train_fixed_ebno_db=None,
                  noise_pilot_fixed=None):
    """One SNR point on this rank's blocks.

    res             Reservoir with n_inputs = 2 N_r, n_outputs = 2 N_t (input_scaling as the template:
                    0.005 / sqrt(var_x), :237-241)
    pil_idx         [G, N, N_t] uint8 pilot symbol indices (one pilot OFDM symbol per coherence block)
    data_idx        [B, N, N_t] uint8 data symbol indices, block_of_frame [B] int32 in [0, G)
    taps            [G, N_r, N_t, isi] complex channel taps
    noise_*         optional standard-normal complex noise [G or B, N+cp, N_r] (else the device stream)
    train_fixed_ebno_db  if set, the template's second ESN is trained as well (pilot re-sent at this Eb/N0 through
                    the same channel with fresh noise, input scaling 0.005 / sqrt(var_x of THAT Eb/N0)) and detects the
                    same data symbols: counters under "ESN_trainFixed"
    Returns {detector: int64 tensor [bit errors, bits]} and the extras used by the tests.
    """
    dev = pil_idx.device
    G, _, N_t = pil_idx.shape
    B = data_idx.shape[0]
    N_r = taps.shape[1]
    cp = isi - 1
    Pi = 10 ** (ebno_db / 10) * No
    var_x = Pi * N
    A_clip = math.sqrt(var_x) * 10 ** (clip_db / 20)
    std = math.sqrt((N + cp) * No / 2)
    if delay is None:                                   # DelayFlag == 0: (Min + Max) // 2, Max = ceil(isi/2) + 2
        delay = (0 + int(math.ceil(isi / 2) + 2)) // 2
    transient = delay + cp
    rd = torch.float64 if fit_precision == "fp64" else torch.float32
    taps = taps.to(dev)
    # Two streams.  The readout-training chain (pilot harvest on a few 16-CTA clusters, Gram, one Cholesky CTA per
    # block) is latency-bound and leaves most SMs idle; everything that does not depend on the readouts -- data
    # frame synthesis, the channel estimates and the three comparison detectors -- runs beside it on a side stream
    # and joins the main stream before the ESN detects (esn_in) and at the end (the counters).
    if detect_precision in ("tc", "auto"):              # decided first, while nothing is queued: a host sync here is free
        # resident tensor-core kernel when every readout owns whole aligned tiles, else (blocks of 18 frames as in the
        # demos, reservoirs above 512 neurons) the streamed-state kernel, which takes any frame -> readout map;
        # 'auto' falls back to the cluster / SIMT kernels for small batches
        detect_precision = res.auto_predict_path(B, block_of_frame) if detect_precision == "auto" else \
            ("tc" if res._tc_resident_ok(None, block_of_frame) else "tcs")
    main = torch.cuda.current_stream(dev)
    side = _side_stream(dev)
    side.wait_stream(main)
    # ---- pilot symbol of every block: ESN training pair through the same channel / noise as the comb pilot
    # (enqueued first, so that the harvest clusters are placed before the side stream fills the other SMs)
    pil = ofdm.synth_frames(pil_idx, taps, Pi, A_clip, N, cp, qam_bits, std, delay=delay, noise=noise_pilot,
                            seed=seed, dtype=torch.float64, want_x_cp=True)
    teacher = torch.zeros((G, N + cp + delay, 2 * N_t), dtype=torch.float64, device=dev)
    teacher[:, delay:, :] = torch.view_as_real(pil["x_cp"]).reshape(G, N + cp, 2 * N_t)
    ext = res.harvest(pil["esn_in"].to(rd), teacher.to(rd), precision=fit_precision, seed=state_noise_seed)
    W_out, info = res.train_readout(ext, teacher, transient)
    dd = torch.float32 if detect_precision in ("fp32", "tc", "tcs") else torch.float64
    W_out_f = res_f = None
    if train_fixed_ebno_db is not None:
        Pi_f = 10 ** (train_fixed_ebno_db / 10) * No
        var_x_f = Pi_f * N
        pil_f = ofdm.synth_frames(pil_idx, taps, Pi_f, math.sqrt(var_x_f) * 10 ** (clip_db / 20), N, cp, qam_bits, std,
                                  delay=delay, noise=noise_pilot_fixed, seed=seed + 2, dtype=torch.float64, want_x_cp=True)
        teacher_f = torch.zeros((G, N + cp + delay, 2 * N_t), dtype=torch.float64, device=dev)
        teacher_f[:, delay:, :] = torch.view_as_real(pil_f["x_cp"]).reshape(G, N + cp, 2 * N_t)
        res_f = res.rescaled(input_scaling=res._aff[1]["in_scale"] * math.sqrt(var_x / var_x_f))
        ext_f = res_f.harvest(pil_f["esn_in"].to(rd), teacher_f.to(rd), precision=fit_precision, seed=state_noise_seed + 2)
        W_out_f, info_f = res_f.train_readout(ext_f, teacher_f, transient)
        info = torch.maximum(info.abs(), info_f.abs())
        del ext_f, pil_f
    out = {}
    with torch.cuda.stream(side):
        fr = ofdm.synth_frames(data_idx, taps.to(torch.complex64 if dd == torch.float32 else torch.complex128), Pi, A_clip,
                               N, cp, qam_bits, std, delay=delay, chan_index=block_of_frame, noise=noise_data,
                               seed=seed + 1, dtype=dd)
        frames_ready = torch.cuda.Event()
        frames_ready.record(side)
        # the input pre-scale of the tensor-core detect (max |u| of the frames, one host sync) off the main stream
        su_exp = res.input_scale_exponent(fr["esn_in"]) if detect_precision == "tc" else None
        # channel estimates from the comb pilot (:316-334) and the true channel
        ls = ofdm.synth_frames(comb_pilot(pil_idx), taps, Pi, A_clip, N, cp, qam_bits, std, noise=noise_pilot,
                               seed=seed, dtype=torch.float64, want_esn_in=False)
        const = const_table(qam_bits, dev)
        X_LS = torch.where(comb_pilot(pil_idx) == 255, torch.zeros((), dtype=const.dtype, device=dev),
                           const[pil_idx.long()])
        Y_LS = ofdm.rx_fft(ls["y_cp"], N, cp)
        H_LS, H_MMSE = ofdm.chanest(Y_LS, X_LS, Pi, isi_profile(isi, dev), isi, No)
        H_true = channel_frequency_response(taps, N)                                                   # [G,N,N_r,N_t]
        Y = ofdm.rx_fft(fr["y_cp"], N, cp)
        cd = Y.dtype
        for name, H, reg in (("Perfect_ZF", H_true, 1e-12), ("LS_ZF", H_LS, 1e-12), ("MMSE", H_MMSE, No / Pi)):
            X = ofdm.equalize(Y, H.to(cd), reg, math.sqrt(Pi), h_index=block_of_frame)
            _, c = ofdm.demap_count(X, qam_bits, tx_idx=data_idx, want_idx=False)
            out[name] = c
            del X
    del ext
    # ---- data symbols
    main.wait_event(frames_ready)
    fr["esn_in"].record_stream(main)
    esn_in, gids = fr["esn_in"], block_of_frame
    total = B * N * N_t * qam_bits

    def esn_detect(r, W, sd):
        if detect_precision == "tc":                    # layout checked above: straight to the kernel, no host sync
            su = su_exp if r is res else r.input_scale_exponent(esn_in)      # (the second reservoir scales its inputs differently)
            if r.tcr_supported():                       # fp32 readout on the CUDA cores
                y = r.predict_tcr(esn_in, r.tcs_prepare(W), transient=transient, group_ids=gids, seed=sd, su_exp=su)   # (layout verdict cached per tensor)
            else:
                y = r.predict_tc(esn_in, r.tc_prepare(W, su), transient=transient, group_ids=gids, seed=sd)
        elif detect_precision == "tcs":
            y = r.predict_tcs(esn_in, r.tcs_prepare(W), transient=transient, group_ids=gids, seed=sd)
        else:
            y = r.predict(esn_in, W, transient=transient, group_ids=gids, precision=detect_precision, seed=sd)
        return ofdm.unpack_fft_demap(y, N, N_t, Pi, qam_bits, tx_idx=data_idx, want_xhat=False, want_idx=False)[2]
    out["ESN"] = esn_detect(res, W_out, state_noise_seed + 1)
    if W_out_f is not None:
        out["ESN_trainFixed"] = esn_detect(res_f, W_out_f, state_noise_seed + 3)
    if int(info.abs().max()) != 0:                      # after the detect is queued: the check costs no GPU idle time
        raise np.linalg.LinAlgError("readout training failed for block %d" % int(torch.nonzero(info)[0]))
    main.wait_stream(side)
    for v in out.values():
        v.record_stream(main)
    tot_t = torch.full((), total, dtype=torch.int64, device=dev)
    res_ = {k: torch.stack([v[0], tot_t.to(v.dtype)]) for k, v in out.items()}
    res_["_W_out"], res_["_delay"], res_["_transient"] = W_out, delay, transient
    if W_out_f is not None:
        res_["_W_out_trainFixed"] = W_out_f
    return res_


def ber_curve(res_factory, N_t, N_r, N, qam_bits, ebno_db_list, n_blocks, frames_per_block, isi=8, No=1e-5, seed=0,
              fit_precision="fp64", detect_precision="tc", device=None, channel="rayleigh", fs_hz=2 * 1.024e6,
              ds_ns=300.0, shard=True, max_blocks_per_launch=None, train_fixed_ebno_db=None):
    """BER-vs-SNR Monte-Carlo: for every Eb/N0, `n_blocks` coherence blocks of `frames_per_block` data
    symbols (blocks sharded over ranks, counters summed over ranks).  `res_factory(var_x)` returns the
    Reservoir for an SNR point (the template scales the inputs by 0.005 / sqrt(var_x)).  Returns
    {detector: [BER per SNR]} plus 'EBN0'.  channel: 'rayleigh' (exponential 8-tap profile of the NBF
    template) or 'tdlb' (the CDL demo's TDL-B taps at sample rate fs_hz, delay spread ds_ns).
    shard=False: this rank runs all blocks by itself and no collective is issued (hyper-parameter sweeps
    place whole configurations on ranks instead, examples/esn_sweep.py); the counters are returned under
    '_counts' ([n_snr, n_detectors, 2] int64) for the caller to gather.
    A rank's blocks are processed `max_blocks_per_launch` at a time (default: ~75 K frames per launch, i.e. 592
    blocks x 128 frames or 4209 blocks x 18 frames, between 148 and 4736 blocks = 10 GB of
    frames and intermediates; the pilots of a launch are one fit batch -- 148 per launch left the fp64 harvest,
    Gram and Cholesky at a quarter of a wave: 475 K frames/s against 660 K), so a point of BASELINE.json configs[4] -- 10^6 frames per Eb/N0 -- runs in
    bounded memory; the counters accumulate on the device."""
    device = device or torch.device("cuda", torch.cuda.current_device())
    if not max_blocks_per_launch:
        max_blocks_per_launch = min(4736, max(148, 75776 // max(1, int(frames_per_block))))
    rank, world = (D.rank(), D.world()) if shard else (0, 1)
    g0, g1 = D.shard_range(n_blocks, rank, world)
    G = g1 - g0
    gen = torch.Generator(device=device)
    dets = DETECTORS_TRAIN_FIXED if train_fixed_ebno_db is not None else DETECTORS
    curves = {k: [] for k in dets}
    all_counts = []
    for si, ebno in enumerate(ebno_db_list):
        gen.manual_seed(seed * 100003 + si * 1009 + rank)
        counts = torch.zeros((len(dets), 2), dtype=torch.int64, device=device)
        res = res_factory(10 ** (ebno / 10) * No * N) if G > 0 else None
        for ci, c0 in enumerate(range(0, G, max(1, int(max_blocks_per_launch)))):
            Gc = min(int(max_blocks_per_launch), G - c0)
            if channel == "tdlb":
                taps = tdlb_taps(Gc, N_r, N_t, isi, fs_hz, ds_ns, gen, device)
            else:                                       # exponential-profile Rayleigh taps of the NBF template (:276-277)
                mag = isi_profile(isi, device)
                taps = (torch.randn((Gc, N_r, N_t, isi), generator=gen, device=device, dtype=torch.float64)
                        + 1j * torch.randn((Gc, N_r, N_t, isi), generator=gen, device=device, dtype=torch.float64)) / math.sqrt(2)
                taps = taps * mag.sqrt()
            pil_idx = torch.randint(0, 2 ** qam_bits, (Gc, N, N_t), generator=gen, device=device, dtype=torch.uint8)
            data_idx = torch.randint(0, 2 ** qam_bits, (Gc * frames_per_block, N, N_t), generator=gen, device=device,
                                     dtype=torch.uint8)
            blk = (torch.arange(Gc * frames_per_block, device=device) // frames_per_block).to(torch.int32)
            r = detect_blocks(res, pil_idx, data_idx, blk, taps, ebno, N, qam_bits, isi=isi, No=No,
                              fit_precision=fit_precision, detect_precision=detect_precision,
                              seed=seed * 7919 + si * 31 + rank * 3 + 11 + ci * 104729,
                              state_noise_seed=seed + 17 * si + rank + ci * 7907,
                              train_fixed_ebno_db=train_fixed_ebno_db)
            for di, k in enumerate(dets):
                counts[di] += r[k]
        if shard:
            D.allreduce_sum_(counts)
        all_counts.append(counts)
        for di, k in enumerate(dets):
            curves[k].append(float(counts[di, 0]) / max(1, int(counts[di, 1])))
    curves["EBN0"] = [int(e) if float(e).is_integer() else float(e) for e in ebno_db_list]
    curves["_counts"] = torch.stack(all_counts)
    return curves
