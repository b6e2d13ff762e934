"""Batched ESN engine on one B200: the host-side mirror of the reference's
`pyESN.ESN` numerics, operating on torch CUDA tensors and calling the C-ABI
kernels (include/esn_b200.h) through ctypes.  torch is used for device memory
and streams only.

Frames are independent: `predict()` never writes the ESN's state back
(reference libs/pyESN.py:218-255), so B frames (and G trained readouts, one per
channel realisation) are stepped in one launch.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import ESN_F32, ESN_F64, MODE_HARVEST, MODE_PREDICT, EsnB200Error, check, ptr

_TORCH = {ESN_F32: torch.float32, ESN_F64: torch.float64}
_CODE = {torch.float32: ESN_F32, torch.float64: ESN_F64, "fp32": ESN_F32, "fp64": ESN_F64,
         "f32": ESN_F32, "f64": ESN_F64}


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _require_cuda():
    if not torch.cuda.is_available():
        raise EsnB200Error("no CUDA device: the esn_b200 engine has no CPU fallback")


def dtype_code(precision):
    try:
        return _CODE[precision]
    except KeyError:
        raise ValueError(f"precision must be 'fp32' or 'fp64', got {precision!r}")


def _vec(v, n, default, device):
    """None / scalar / length-n vector (numpy or torch, any device) -> fp64 device vector of length n."""
    if isinstance(v, torch.Tensor):
        t = v.detach().to(device=device, dtype=torch.float64).reshape(-1)
        return (t.expand(n) if t.numel() == 1 else t.reshape(n)).contiguous()
    if v is None:
        a = np.full(n, default, dtype=np.float64)
    else:
        a = np.asarray(v, dtype=np.float64)
        if a.ndim == 0:
            a = np.full(n, float(a), dtype=np.float64)
        a = np.ascontiguousarray(a.reshape(n))
    return torch.from_numpy(a).to(device)


class TcReadout:
    """Handle of the tensor-core weight image built by Reservoir.tc_prepare."""

    def __init__(self, weights, image, yscale, su_exp, sy_exp, n_groups):
        self.weights, self.image, self.yscale = weights, image, yscale
        self.su_exp, self.sy_exp, self.n_groups = su_exp, sy_exp, n_groups


class TcsReadout:
    """fp32 readout tables of the streamed-state tensor-core kernel (Reservoir.tcs_prepare)."""

    def __init__(self, wo_x, wo_u, n_groups):
        self.wo_x, self.wo_u, self.n_groups = wo_x, wo_u, n_groups


class Reservoir:
    """Device-resident reservoir weights and the affine I/O maps of one ESN
    (reference libs/pyESN.py:93-152).  Weights are uploaded once, in the
    augmented transposed layout the kernels stream:
    Wt_aug[K_aug_pad][N_pad] = [W^T; W_in^T; W_fb^T; 0]."""

    def __init__(self, W, W_in, W_fb, input_scaling=None, input_shift=None,
                 teacher_scaling=None, teacher_shift=None, noise=0.0,
                 teacher_forcing=True, device="cuda"):
        _require_cuda()
        self.lib = _lib.load()
        self.device = torch.device(device)
        W = np.asarray(W, dtype=np.float64)
        W_in = np.asarray(W_in, dtype=np.float64)
        W_fb = np.asarray(W_fb, dtype=np.float64)
        self.N, self.n_in, self.n_out = W.shape[0], W_in.shape[1], W_fb.shape[1]
        if self.n_out > _lib.ESN_MAX_OUT or self.n_in > _lib.ESN_MAX_IN:
            raise EsnB200Error(f"n_inputs <= {_lib.ESN_MAX_IN} and n_outputs <= {_lib.ESN_MAX_OUT} required")
        self.P = self.N + self.n_in
        npad, kpad = C.c_int(), C.c_int()
        check(self.lib.esn_pad_sizes(self.N, self.n_in, self.n_out, C.byref(npad), C.byref(kpad)), "esn_pad_sizes")
        self.N_pad, self.K_aug_pad = npad.value, kpad.value
        self.noise = float(noise)
        self.teacher_forcing = bool(teacher_forcing)
        wt = np.zeros((self.K_aug_pad, self.N_pad), dtype=np.float64)
        wt[:self.N, :self.N] = W.T
        wt[self.N:self.P, :self.N] = W_in.T
        if self.teacher_forcing:
            wt[self.P:self.P + self.n_out, :self.N] = W_fb.T
        # plain fp64 copies for the tensor-core path's weight folding (esn_tc_prepare)
        self._W64 = torch.from_numpy(np.ascontiguousarray(W)).to(self.device)
        self._Win64 = torch.from_numpy(np.ascontiguousarray(W_in)).to(self.device)
        self._Wfb64 = torch.from_numpy(np.ascontiguousarray(W_fb)).to(self.device)
        self._tc_images = {}
        self._wt = {ESN_F64: torch.from_numpy(wt).to(self.device)}
        self._wt[ESN_F32] = self._wt[ESN_F64].to(torch.float32)
        self._aff = {ESN_F64: dict(
            in_scale=_vec(input_scaling, self.n_in, 1.0, self.device),
            in_shift=_vec(input_shift, self.n_in, 0.0, self.device),
            t_scale=_vec(teacher_scaling, self.n_out, 1.0, self.device),
            t_shift=_vec(teacher_shift, self.n_out, 0.0, self.device))}
        self._aff[ESN_F32] = {k: v.to(torch.float32) for k, v in self._aff[ESN_F64].items()}

    def rescaled(self, input_scaling=None, input_shift=None, teacher_scaling=None, teacher_shift=None, noise=None):
        """The same reservoir with other affine I/O maps (the demos rescale the inputs by 0.005 / sqrt(var_x) at
        every Eb/N0, OFDM_MIMO_2-2_NBF_LDPC.py:237-241): shares the device copies of the weights and the
        tensor-core weight images, so nothing is uploaded or rebuilt.  Arguments left at None are kept."""
        import copy
        other = copy.copy(self)
        cur = self._aff[ESN_F64]
        if "_y_absmax" not in cur:
            cur["_y_absmax"] = float((cur["t_scale"].abs() * 4 + cur["t_shift"].abs()).max().item())
        other._aff = {ESN_F64: dict(
            in_scale=cur["in_scale"] if input_scaling is None else _vec(input_scaling, self.n_in, 1.0, self.device),
            in_shift=cur["in_shift"] if input_shift is None else _vec(input_shift, self.n_in, 0.0, self.device),
            t_scale=cur["t_scale"] if teacher_scaling is None else _vec(teacher_scaling, self.n_out, 1.0, self.device),
            t_shift=cur["t_shift"] if teacher_shift is None else _vec(teacher_shift, self.n_out, 0.0, self.device))}
        if teacher_scaling is None and teacher_shift is None:
            other._aff[ESN_F64]["_y_absmax"] = cur["_y_absmax"]
        other._aff[ESN_F32] = {k: v.to(torch.float32) for k, v in other._aff[ESN_F64].items() if k != "_y_absmax"}
        if noise is not None:
            other.noise = float(noise)
        return other

    def with_affine(self, input_scaling, input_shift, teacher_scaling, teacher_shift, noise):
        """The same reservoir with ALL affine maps and the noise amplitude replaced (None = the reference's
        "no scaling / no shift"); shares the weight uploads and tensor-core images."""
        import copy
        other = copy.copy(self)
        other._aff = {ESN_F64: dict(in_scale=_vec(input_scaling, self.n_in, 1.0, self.device),
                                    in_shift=_vec(input_shift, self.n_in, 0.0, self.device),
                                    t_scale=_vec(teacher_scaling, self.n_out, 1.0, self.device),
                                    t_shift=_vec(teacher_shift, self.n_out, 0.0, self.device))}
        other._aff[ESN_F32] = {k: v.to(torch.float32) for k, v in other._aff[ESN_F64].items()}
        other.noise = float(noise)
        return other

    @staticmethod
    def _check_group_ids(group_ids, n_groups):
        """Host-resident ids are validated for free; device-resident ones are clamped by the kernels."""
        if isinstance(group_ids, np.ndarray) or (isinstance(group_ids, torch.Tensor) and not group_ids.is_cuda):
            g = np.asarray(group_ids)
            if g.size and (g.min() < 0 or g.max() >= n_groups):
                raise ValueError(f"group_ids must lie in [0, {n_groups}), got [{g.min()}, {g.max()}]")

    # ------------------------------------------------------------------ run --
    def _run(self, mode, code, inputs, teachers=None, W_out=None, group_ids=None, x0=None,
             y0=None, noise_uniforms=None, seed=0, transient=0, want_ext=False):
        td = _TORCH[code]
        inputs = self._as(inputs, td, 3)
        B, T, n_in = inputs.shape
        if n_in != self.n_in:
            raise ValueError(f"inputs have {n_in} columns, ESN has n_inputs={self.n_in}")
        a = _lib.RecurrenceArgs()
        a.dtype, a.mode, a.B, a.T = code, mode, B, T
        a.N, a.n_in, a.n_out = self.N, self.n_in, self.n_out
        a.N_pad, a.K_aug_pad = self.N_pad, self.K_aug_pad
        a.transient, a.feedback = int(transient), int(self.teacher_forcing)
        a.noise_amp, a.seed = self.noise, int(seed) & 0xFFFFFFFFFFFFFFFF
        aff = self._aff[code]
        keep = [inputs]
        a.Wt_aug, a.inp = ptr(self._wt[code]), ptr(inputs)
        a.in_scale, a.in_shift = ptr(aff["in_scale"]), ptr(aff["in_shift"])
        a.t_scale, a.t_shift = ptr(aff["t_scale"]), ptr(aff["t_shift"])
        steps = T if mode == MODE_PREDICT else T - 1
        if noise_uniforms is not None:
            noise_uniforms = self._as(noise_uniforms, td, 3)
            if tuple(noise_uniforms.shape) != (B, steps, self.N):
                raise ValueError(f"noise_uniforms must be [{B},{steps},{self.N}]")
            a.noise_uniforms = ptr(noise_uniforms)
            keep.append(noise_uniforms)
        ext = None
        if mode == MODE_HARVEST or want_ext:
            ext = torch.empty((B, T, self.P), dtype=td, device=self.device)
            a.ext_out = ptr(ext)
        ws = torch.empty((B, self.N), dtype=td, device=self.device)
        a.workspace = ptr(ws)
        y = None
        if mode == MODE_HARVEST:
            teachers = self._as(teachers, td, 3)
            if tuple(teachers.shape) != (B, T, self.n_out):
                raise ValueError(f"teachers must be [{B},{T},{self.n_out}]")
            a.teacher = ptr(teachers)
            keep.append(teachers)
        else:
            W_out = self._as(W_out, td, 3)
            if W_out.shape[1:] != (self.n_out, self.P):
                raise ValueError(f"W_out must be [G,{self.n_out},{self.P}]")
            a.W_out, a.n_groups = ptr(W_out), W_out.shape[0]
            keep.append(W_out)
            if group_ids is not None:
                self._check_group_ids(group_ids, W_out.shape[0])
                group_ids = torch.as_tensor(group_ids).to(device=self.device, dtype=torch.int32).contiguous()
                a.group_ids = ptr(group_ids)
                keep.append(group_ids)
            elif W_out.shape[0] != 1:
                raise ValueError("group_ids required when W_out holds more than one readout")
            if x0 is not None:
                x0 = self._as(x0, td, 2)
                a.x0 = ptr(x0)
                keep.append(x0)
            if y0 is not None:
                y0 = self._as(y0, td, 2)
                a.y0 = ptr(y0)
                keep.append(y0)
            y = torch.empty((B, T - int(transient), self.n_out), dtype=td, device=self.device)
            a.y_out = ptr(y)
        check(self.lib.esn_recurrence_run(C.byref(a), _stream()), "esn_recurrence_run")
        return ext, y

    def _as(self, t, td, ndim):
        if isinstance(t, np.ndarray):
            t = torch.from_numpy(np.ascontiguousarray(t))
        t = t.to(device=self.device, dtype=td)
        while t.dim() < ndim:
            t = t.unsqueeze(0)
        return t.contiguous()

    # ---------------------------------------------------- tensor-core path --
    def tc_supported(self):
        return bool(self.lib.esn_tc_supported(self.N, self.n_in, self.n_out))

    def tc_tile_frames(self):
        """Frames that must share one readout on the tensor-core path: the 64 frames of one CTA of a pair when
        the ESN has at most 8 outputs (the 16 readout rows of the MMA hold one readout per CTA), else the 128
        frames of the pair."""
        return 64 if self.n_out <= 8 else 128

    def input_scale_exponent(self, inputs):
        """su_exp for the tensor-core path: the power of two that brings the
        largest scaled input to about 2^9 (one device reduction + sync)."""
        aff = self._aff[ESN_F32]
        u = inputs.to(self.device, torch.float32) * aff["in_scale"] + aff["in_shift"]
        umax = float(u.abs().max().item())
        if not np.isfinite(umax):
            raise EsnB200Error("non-finite inputs")
        e = int(np.ceil(np.log2(umax))) if umax > 0 else 0
        su = 9 - e
        if su < 1:
            raise EsnB200Error("inputs too large for the tensor-core path (max |u| > 256); use precision='fp32'")
        return min(su, 24)

    def output_scale_exponent(self, y_absmax):
        """sy_exp for the tensor-core path: brings the largest fed-back output
        (scaled teacher domain) to about 2^6, leaving 2^10 of fp16 headroom."""
        y_absmax = float(y_absmax)
        if not np.isfinite(y_absmax) or y_absmax <= 0:
            return 0
        return int(np.clip(6 - int(np.ceil(np.log2(y_absmax))), -8, 40))

    def _tc_weights(self, su_exp, sy_exp):
        key = (int(su_exp), int(sy_exp))
        if key not in self._tc_images:
            nbytes = int(self.lib.esn_tc_weight_bytes(self.N, self.n_in))
            image = torch.empty((nbytes,), dtype=torch.uint8, device=self.device)
            check(self.lib.esn_tc_prepare_weights(ptr(self._W64), ptr(self._Win64), ptr(self._Wfb64), self.N,
                                                  self.n_in, self.n_out, key[0], key[1],
                                                  int(self.teacher_forcing), ptr(image), _stream()),
                  "esn_tc_prepare_weights")
            self._tc_images[key] = image
        return self._tc_images[key]

    def tc_prepare(self, W_out, su_exp, y_absmax=None):
        """Build the per-readout tensor-core images for W_out [G, n_out, P]
        (and, once per (su, sy), the shared weight image).  `y_absmax`: largest
        |scaled teacher| the readouts were trained on (sets the feedback
        pre-scale); defaults to the t_scale magnitude."""
        if not self.tc_supported():
            raise EsnB200Error("tensor-core path needs N <= 512, n_inputs <= 24, n_outputs <= 16")
        W_out = self._as(W_out, torch.float64, 3)
        G = W_out.shape[0]
        if y_absmax is None:                               # cached per teacher map: a device reduction + sync here
            aff = self._aff[ESN_F64]                       # would stall the thread that is enqueueing the detect
            if "_y_absmax" not in aff:
                aff["_y_absmax"] = float((aff["t_scale"].abs() * 4 + aff["t_shift"].abs()).max().item())
            y_absmax = aff["_y_absmax"]
        sy_exp = self.output_scale_exponent(y_absmax)
        weights = self._tc_weights(su_exp, sy_exp)
        nbytes = int(self.lib.esn_tc_readout_bytes(self.N, self.n_in))
        image = torch.empty((G, nbytes), dtype=torch.uint8, device=self.device)
        yscale = torch.empty((G,), dtype=torch.float32, device=self.device)
        check(self.lib.esn_tc_prepare_readout(ptr(W_out), self.N, self.n_in, self.n_out, G, int(su_exp),
                                              ptr(image), ptr(yscale), _stream()), "esn_tc_prepare_readout")
        return TcReadout(weights, image, yscale, int(su_exp), sy_exp, G)

    def predict_tc(self, inputs, readout, transient=0, group_ids=None, x0=None, y0=None,
                   noise_uniforms=None, seed=0, return_ext=False, timeline=None):
        """Free-running prediction on the tensor cores.  `readout` comes from
        tc_prepare; each 128-frame tile must use a single readout."""
        inputs = self._as(inputs, torch.float32, 3)
        B, T, n_in = inputs.shape
        if n_in != self.n_in:
            raise ValueError(f"inputs have {n_in} columns, ESN has n_inputs={self.n_in}")
        a = _lib.TcPredictArgs()
        a.B, a.T, a.N, a.n_in, a.n_out = B, T, self.N, self.n_in, self.n_out
        a.transient, a.feedback = int(transient), int(self.teacher_forcing)
        a.su_exp, a.sy_exp, a.n_groups = readout.su_exp, readout.sy_exp, readout.n_groups
        a.noise_amp, a.seed = self.noise, int(seed) & 0xFFFFFFFFFFFFFFFF
        aff = self._aff[ESN_F32]
        a.weights, a.readouts, a.yscale = ptr(readout.weights), ptr(readout.image), ptr(readout.yscale)
        a.inp = ptr(inputs)
        a.in_scale, a.in_shift = ptr(aff["in_scale"]), ptr(aff["in_shift"])
        a.t_scale, a.t_shift = ptr(aff["t_scale"]), ptr(aff["t_shift"])
        keep = [inputs]
        if group_ids is not None:
            self._check_group_ids(group_ids, readout.n_groups)
            group_ids = torch.as_tensor(group_ids).to(device=self.device, dtype=torch.int32).contiguous()
            a.group_ids = ptr(group_ids)
            keep.append(group_ids)
        elif readout.n_groups != 1:
            raise ValueError("group_ids required when the readout handle holds more than one readout")
        if x0 is not None:
            x0 = self._as(x0, torch.float32, 2)
            a.x0 = ptr(x0)
        if y0 is not None:
            y0 = self._as(y0, torch.float32, 2)
            a.y0 = ptr(y0)
        if noise_uniforms is not None:
            noise_uniforms = self._as(noise_uniforms, torch.float32, 3)
            if tuple(noise_uniforms.shape) != (B, T, self.N):
                raise ValueError(f"noise_uniforms must be [{B},{T},{self.N}]")
            a.noise_uniforms = ptr(noise_uniforms)
        ext = None
        if return_ext:
            ext = torch.empty((B, T, self.P), dtype=torch.float32, device=self.device)
            a.ext_out = ptr(ext)
        y = torch.empty((B, T - int(transient), self.n_out), dtype=torch.float32, device=self.device)
        a.y_out = ptr(y)
        if timeline is not None:           # [T+1, 8] int64 device tensor (profiling aid)
            a.timeline = ptr(timeline)
        check(self.lib.esn_tc_predict(C.byref(a), _stream()), "esn_tc_predict")
        return (y, ext) if return_ext else y

    def harvest_tc(self, inputs, teachers, noise_uniforms=None, seed=0):
        """Teacher-forced harvesting on the tensor cores (CTA-pair kernel, N padded to
        256 or 512).  Throughput mode: states carry the tensor-core path's ~1e-5
        relative error, so W_out trained on them differs from the fp64 fit by more than
        the 1e-4 parity bar (the noise regulariser itself moves states by ~1e-2); use
        precision='fp64' where W_out parity is asserted."""
        if not self.tc_supported():
            raise EsnB200Error("tensor-core path needs N <= 512, n_inputs <= 24, n_outputs <= 16")
        inputs = self._as(inputs, torch.float32, 3)
        teachers = self._as(teachers, torch.float32, 3)
        B, T, n_in = inputs.shape
        if n_in != self.n_in or tuple(teachers.shape) != (B, T, self.n_out):
            raise ValueError(f"inputs [B,T,{self.n_in}] and teachers [B,T,{self.n_out}] expected")
        aff = self._aff[ESN_F32]
        y_absmax = float((teachers * aff["t_scale"] + aff["t_shift"]).abs().max().item())
        su_exp, sy_exp = self.input_scale_exponent(inputs), self.output_scale_exponent(y_absmax)
        a = _lib.TcPredictArgs()
        a.B, a.T, a.N, a.n_in, a.n_out = B, T, self.N, self.n_in, self.n_out
        a.transient, a.feedback = 0, int(self.teacher_forcing)
        a.su_exp, a.sy_exp, a.n_groups = su_exp, sy_exp, 0
        a.noise_amp, a.seed = self.noise, int(seed) & 0xFFFFFFFFFFFFFFFF
        weights = self._tc_weights(su_exp, sy_exp)
        a.weights, a.inp, a.teacher = ptr(weights), ptr(inputs), ptr(teachers)
        a.in_scale, a.in_shift = ptr(aff["in_scale"]), ptr(aff["in_shift"])
        a.t_scale, a.t_shift = ptr(aff["t_scale"]), ptr(aff["t_shift"])
        if noise_uniforms is not None:
            noise_uniforms = self._as(noise_uniforms, torch.float32, 3)
            if tuple(noise_uniforms.shape) != (B, T - 1, self.N):
                raise ValueError(f"noise_uniforms must be [{B},{T - 1},{self.N}]")
            a.noise_uniforms = ptr(noise_uniforms)
        ext = torch.empty((B, T, self.P), dtype=torch.float32, device=self.device)
        a.ext_out = ptr(ext)
        check(self.lib.esn_tc_predict(C.byref(a), _stream()), "esn_tc_predict(harvest)")
        return ext

    # ------------------------------------- tensor cores, state streamed through L2 --
    def tcs_supported(self):
        return bool(self.lib.esn_tcs_supported(self.N, self.n_in, self.n_out))

    def tcs_prepare(self, W_out):
        """fp32 readout tables for W_out [G, n_out, P] (any G; frames pick their readout freely)."""
        W_out = self._as(W_out, torch.float64, 3)
        G = W_out.shape[0]
        nu = C.c_longlong()
        nx = int(self.lib.esn_tcs_readout_floats(self.N, self.n_out, C.byref(nu)))
        wo_x = torch.empty((G, nx), dtype=torch.float32, device=self.device)
        wo_u = torch.empty((G, nu.value), dtype=torch.float32, device=self.device)
        check(self.lib.esn_tcs_prepare_readout(ptr(W_out), self.N, self.n_in, self.n_out, G, ptr(wo_x), ptr(wo_u),
                                               _stream()), "esn_tcs_prepare_readout")
        return TcsReadout(wo_x, wo_u, G)

    def _tcs_workspace(self, B):
        nbytes = int(self.lib.esn_tcs_workspace_bytes(B, self.N))
        ws = getattr(self, "_tcs_ws", None)
        if ws is None or ws.numel() < nbytes or ws.device != self.device:
            ws = torch.empty((nbytes,), dtype=torch.uint8, device=self.device)
            self._tcs_ws = ws
        return ws

    def _tcs_run(self, inputs, readout=None, teachers=None, transient=0, group_ids=None, x0=None, y0=None,
                 noise_uniforms=None, seed=0, return_ext=False, y_absmax=None, tune=None, timeline=None,
                 resident=False, su_exp=None):
        if resident and not self.tcr_supported():
            raise EsnB200Error("resident tensor-core path needs N <= 512, n_inputs <= 16, n_outputs <= 8")
        if not self.tcs_supported():
            raise EsnB200Error("tensor-core path needs N <= 4096, n_inputs <= 24, n_outputs <= 16")
        inputs = self._as(inputs, torch.float32, 3)
        B, T, n_in = inputs.shape
        if n_in != self.n_in:
            raise ValueError(f"inputs have {n_in} columns, ESN has n_inputs={self.n_in}")
        harvest = teachers is not None
        aff = self._aff[ESN_F32]
        if harvest:
            teachers = self._as(teachers, torch.float32, 3)
            if tuple(teachers.shape) != (B, T, self.n_out):
                raise ValueError(f"teachers must be [{B},{T},{self.n_out}]")
            if y_absmax is None:               # (a device reduction + sync; callers in a pipeline pass it)
                y_absmax = float((teachers * aff["t_scale"] + aff["t_shift"]).abs().max().item())
        elif y_absmax is None:
            a64 = self._aff[ESN_F64]
            if "_y_absmax" not in a64:
                a64["_y_absmax"] = float((a64["t_scale"].abs() * 4 + a64["t_shift"].abs()).max().item())
            y_absmax = a64["_y_absmax"]
        # su_exp: the caller may pass the input pre-scale exponent (input_scale_exponent costs a device reduction + sync)
        su_exp = self.input_scale_exponent(inputs) if su_exp is None else int(su_exp)
        sy_exp = self.output_scale_exponent(y_absmax)
        a = _lib.TcsArgs()
        a.B, a.T, a.N, a.n_in, a.n_out = B, T, self.N, self.n_in, self.n_out
        a.transient, a.feedback = int(transient), int(self.teacher_forcing)
        a.su_exp, a.sy_exp = su_exp, sy_exp
        a.noise_amp, a.seed = self.noise, int(seed) & 0xFFFFFFFFFFFFFFFF
        if tune:
            a.accumulators, a.ring_a, a.ring_b = (int(tune.get(k, 0)) for k in ("accumulators", "ring_a", "ring_b"))
        weights = self._tc_weights(su_exp, sy_exp)
        ws = None if resident else self._tcs_workspace(B)
        a.weights, a.inp, a.workspace = ptr(weights), ptr(inputs), ptr(ws)
        a.in_scale, a.in_shift = ptr(aff["in_scale"]), ptr(aff["in_shift"])
        a.t_scale, a.t_shift = ptr(aff["t_scale"]), ptr(aff["t_shift"])
        steps = T - 1 if harvest else T
        keep = [inputs, weights, ws]
        if noise_uniforms is not None:
            noise_uniforms = self._as(noise_uniforms, torch.float32, 3)
            if tuple(noise_uniforms.shape) != (B, steps, self.N):
                raise ValueError(f"noise_uniforms must be [{B},{steps},{self.N}]")
            a.noise_uniforms = ptr(noise_uniforms)
        ext = y = None
        if harvest or return_ext:
            ext = torch.empty((B, T, self.P), dtype=torch.float32, device=self.device)
            a.ext_out = ptr(ext)
        if harvest:
            a.teacher = ptr(teachers)
        else:
            a.wo_x, a.wo_u, a.n_groups = ptr(readout.wo_x), ptr(readout.wo_u), readout.n_groups
            if group_ids is not None:
                self._check_group_ids(group_ids, readout.n_groups)
                group_ids = torch.as_tensor(group_ids).to(device=self.device, dtype=torch.int32).contiguous()
                a.group_ids = ptr(group_ids)
            elif readout.n_groups != 1:
                raise ValueError("group_ids required when the readout handle holds more than one readout")
            if x0 is not None:
                x0 = self._as(x0, torch.float32, 2)
                a.x0 = ptr(x0)
            if y0 is not None:
                y0 = self._as(y0, torch.float32, 2)
                a.y0 = ptr(y0)
            y = torch.empty((B, T - int(transient), self.n_out), dtype=torch.float32, device=self.device)
            a.y_out = ptr(y)
        if timeline is not None:
            a.timeline = ptr(timeline)
        if resident:
            check(self.lib.esn_tcr_run(C.byref(a), _stream()), "esn_tcr_run")
        else:
            check(self.lib.esn_tcs_run(C.byref(a), _stream()), "esn_tcs_run")
        return y, ext

    def tcr_supported(self):
        return bool(self.lib.esn_tcr_supported(self.N, self.n_in, self.n_out))

    def predict_tcr(self, inputs, readout, transient=0, group_ids=None, x0=None, y0=None, noise_uniforms=None,
                    seed=0, return_ext=False, y_absmax=None, timeline=None, su_exp=None):
        """Free-running prediction on the tensor cores with the state resident in shared memory and the readout
        on the CUDA cores in fp32: reservoirs of up to 512 neurons, n_inputs <= 16, n_outputs <= 8, one readout
        per aligned run of 64 frames.  `readout`: W_out [G, n_out, P] or the handle from tcs_prepare."""
        if not isinstance(readout, TcsReadout):
            readout = self.tcs_prepare(readout)
        if group_ids is not None and not self._tc_resident_ok(inputs, group_ids):
            raise EsnB200Error("resident tensor-core path: each aligned run of 64 frames must share one readout "
                               "(use predict_tcs / precision='tcs' for a free frame -> readout map)")
        y, ext = self._tcs_run(inputs, readout=readout, transient=transient, group_ids=group_ids, x0=x0, y0=y0,
                               noise_uniforms=noise_uniforms, seed=seed, return_ext=return_ext, y_absmax=y_absmax,
                               timeline=timeline, resident=True, su_exp=su_exp)
        return (y, ext) if return_ext else y

    def harvest_tcr(self, inputs, teachers, noise_uniforms=None, seed=0, su_exp=None, y_absmax=None):
        """Teacher-forced harvesting with the resident tensor-core kernel (split accumulators).  su_exp / y_absmax:
        the input pre-scale exponent and the largest |scaled teacher|, if the caller knows them (each costs a device
        reduction and a host sync otherwise)."""
        return self._tcs_run(inputs, teachers=teachers, noise_uniforms=noise_uniforms, seed=seed, resident=True,
                             su_exp=su_exp, y_absmax=y_absmax)[1]

    def predict_tcs(self, inputs, readout, transient=0, group_ids=None, x0=None, y0=None, noise_uniforms=None,
                    seed=0, return_ext=False, y_absmax=None, tune=None, timeline=None, su_exp=None):
        """Free-running prediction on the tensor cores with the state streamed through L2: any reservoir size up to
        4096 neurons, any frame -> readout map.  `readout`: W_out [G, n_out, P] or the handle from tcs_prepare."""
        if not isinstance(readout, TcsReadout):
            readout = self.tcs_prepare(readout)
        y, ext = self._tcs_run(inputs, readout=readout, transient=transient, group_ids=group_ids, x0=x0, y0=y0,
                               noise_uniforms=noise_uniforms, seed=seed, return_ext=return_ext, y_absmax=y_absmax,
                               tune=tune, timeline=timeline, su_exp=su_exp)
        return (y, ext) if return_ext else y

    def harvest_tcs(self, inputs, teachers, noise_uniforms=None, seed=0, tune=None, su_exp=None, y_absmax=None):
        """Teacher-forced harvesting with the streamed-state tensor-core kernel (throughput mode, see harvest_tc)."""
        return self._tcs_run(inputs, teachers=teachers, noise_uniforms=noise_uniforms, seed=seed, tune=tune,
                             su_exp=su_exp, y_absmax=y_absmax)[1]

    def harvest(self, inputs, teachers, precision="fp64", noise_uniforms=None, seed=0, su_exp=None, y_absmax=None):
        """Teacher-forced harvesting (libs/pyESN.py:179-182).  Returns the
        extended states E [B, T, N+n_in] = [x_n, u_n] (libs/pyESN.py:189).  precision 'auto' = 'fp64': the readout
        solve amplifies state errors, and only the fp64 harvest keeps W_out inside the 1e-4 bar with margin
        (measured at cfg3: fp64 3e-9, fp32 1.0e-4, tensor cores 1.9e-4); 'tc' is the throughput mode."""
        if precision == "auto":
            precision = "fp64"
        if precision == "tc":
            if self.tcr_supported():
                return self.harvest_tcr(inputs, teachers, noise_uniforms=noise_uniforms, seed=seed, su_exp=su_exp,
                                        y_absmax=y_absmax)
            if not self.tc_supported():
                return self.harvest_tcs(inputs, teachers, noise_uniforms=noise_uniforms, seed=seed, su_exp=su_exp,
                                        y_absmax=y_absmax)
            return self.harvest_tc(inputs, teachers, noise_uniforms=noise_uniforms, seed=seed)
        if precision == "tc2":
            return self.harvest_tc(inputs, teachers, noise_uniforms=noise_uniforms, seed=seed)
        if precision == "tcs":
            return self.harvest_tcs(inputs, teachers, noise_uniforms=noise_uniforms, seed=seed)
        if precision == "tcr":
            return self.harvest_tcr(inputs, teachers, noise_uniforms=noise_uniforms, seed=seed)
        ext, _ = self._run(MODE_HARVEST, dtype_code(precision), inputs, teachers=teachers,
                           noise_uniforms=noise_uniforms, seed=seed)
        return ext

    def predict(self, inputs, W_out, transient=0, group_ids=None, x0=None, y0=None,
                precision="auto", noise_uniforms=None, seed=0, return_ext=False):
        """Free-running prediction (libs/pyESN.py:243-255) of B frames, frame b
        using readout W_out[group_ids[b]].  Returns y [B, T-transient, n_out]
        in teacher units (and E if return_ext).  precision: 'auto' (the fastest kernel for this batch, reservoir
        and readout layout: auto_predict_path), 'fp64' / 'fp32' (cluster or streaming SIMT kernels), 'tc' (tensor
        cores: a resident kernel, or the streamed-state kernel where those cannot go), 'tcr' / 'tc2' / 'tcs' (one
        tensor-core kernel explicitly)."""
        if precision == "auto":
            if isinstance(W_out, TcReadout):
                precision = "tc"
            elif isinstance(W_out, TcsReadout):
                precision = "tcs"
            else:
                B = inputs.shape[0] if getattr(inputs, "ndim", 3) == 3 else 1
                precision = self.auto_predict_path(B, group_ids)
        # 'tc': the resident kernel with the fp32 CUDA-core readout (esn_recur_tcr) where it can go -- at most 512
        # neurons, 16 inputs, 8 outputs, one readout per aligned run of 64 frames -- else the first resident kernel
        # (readout inside the MMA: wider I/O, throughput mode), else the streamed-state kernel.  'tc2' asks for
        # the first resident kernel explicitly; a TcReadout handle (tc_prepare) implies it.
        if precision == "tc" and isinstance(W_out, TcReadout):
            precision = "tc2"
        if precision == "tc" and isinstance(W_out, TcsReadout):
            precision = "tcr" if (self.tcr_supported() and self._tc_resident_ok(inputs, group_ids)) else "tcs"
        if precision == "tc":
            if not self._tc_resident_ok(inputs, group_ids):
                precision = "tcs"
            else:
                precision = "tcr" if self.tcr_supported() else "tc2"
        if precision == "tcr":
            return self.predict_tcr(inputs, W_out, transient=transient, group_ids=group_ids, x0=x0, y0=y0,
                                    noise_uniforms=noise_uniforms, seed=seed, return_ext=return_ext)
        if precision == "tcs":
            return self.predict_tcs(inputs, W_out, transient=transient, group_ids=group_ids, x0=x0, y0=y0,
                                    noise_uniforms=noise_uniforms, seed=seed, return_ext=return_ext)
        if precision == "tc2":
            inputs = self._as(inputs, torch.float32, 3)
            if not isinstance(W_out, TcReadout):
                W_out = self.tc_prepare(W_out, self.input_scale_exponent(inputs))
            if group_ids is not None and not self._tc_resident_ok(inputs, group_ids):
                raise EsnB200Error(f"tensor-core path: each {self.tc_tile_frames()}-frame tile must share one readout")
            return self.predict_tc(inputs, W_out, transient=transient, group_ids=group_ids, x0=x0, y0=y0,
                                   noise_uniforms=noise_uniforms, seed=seed, return_ext=return_ext)
        ext, y = self._run(MODE_PREDICT, dtype_code(precision), inputs, W_out=W_out,
                           group_ids=group_ids, x0=x0, y0=y0, noise_uniforms=noise_uniforms,
                           seed=seed, transient=transient, want_ext=return_ext)
        return (y, ext) if return_ext else y

    # ---- automatic path selection (north_star: "tcgen05 tiles at large batch, a warp-shuffle FFMA path at small
    # batch, crossover picked from ncu tensor-pipe and HBM counters").  A tensor-core launch costs one tile time
    # whatever the batch (a CTA pair steps 128 frames through all T steps; 74 pairs are resident), the cluster /
    # streaming SIMT kernels cost one wave per ~48-300 frames: below AUTO_TC_MIN_FRAMES[N_pad] frames the SIMT
    # side is faster AND more accurate (4e-7 vs 9e-6 state error).  Table measured by profiles/crossover.py
    # (profiles/r2_crossover.txt, with the ncu counters of both sides at the boundary).
    # (largest N, first batch size at which the tensor-core launch beats the cluster / streaming SIMT kernels):
    # measured 1.30 ms per ~290-frame wave vs 3.3 ms at 100 neurons, 1.77 ms per ~96 frames vs 6.5 ms at 300,
    # 2.55 ms per 48 frames vs 6.1 ms at 512 (profiles/r2_crossover.txt)
    AUTO_TC_MIN_FRAMES = ((128, 640), (256, 448), (384, 320), (512, 100))
    AUTO_TCS_MIN_FRAMES = 1               # above 512 neurons the streaming SIMT kernel never wins (92 ms vs 32 ms at 1024)

    def auto_predict_path(self, B, group_ids=None):
        """'tc' (resident tensor-core kernel), 'tcs' (streamed-state tensor-core kernel) or 'fp32' (cluster kernel
        for small batches / streaming SIMT kernel) for a predict call of B frames."""
        if not self.tcs_supported():
            return "fp32"
        if self.N > 512:
            return "tcs" if B >= self.AUTO_TCS_MIN_FRAMES else "fp32"
        b_min = next(b for n, b in self.AUTO_TC_MIN_FRAMES if self.N <= n)
        if B < b_min:
            return "fp32"
        return "tc" if self._tc_resident_ok(None, group_ids) else "tcs"

    def _tc_resident_ok(self, inputs, group_ids):
        """True when the resident tensor-core kernel can take the call: N <= 512 and every readout shared by whole
        aligned tiles.  Host-resident group ids are inspected for free; device-resident ones cost one sync."""
        if not self.tc_supported():
            return False
        if group_ids is None:
            return True
        # the answer for a device tensor costs a sync: remembered per tensor OBJECT (weak reference + version counter)
        cache = getattr(self, "_tile_ok_cache", None)
        if (isinstance(group_ids, torch.Tensor) and cache is not None and cache[0]() is group_ids
                and cache[1] == group_ids._version):
            return cache[2]
        tiles = torch.as_tensor(group_ids).reshape(-1)
        tile = self.tc_tile_frames()
        pad = (-tiles.numel()) % tile
        if pad:
            tiles = torch.cat([tiles, tiles[-1:].expand(pad)])
        tiles = tiles.view(-1, tile)
        ok = bool((tiles == tiles[:, :1]).all())
        if isinstance(group_ids, torch.Tensor):
            import weakref
            self._tile_ok_cache = (weakref.ref(group_ids), group_ids._version, ok)
        return ok

    # -------------------------------------------------------------- readout --
    # pivot ratio min d_jj / max d_jj below which the lambda = 0 normal equations are not trusted.  Calibrated on
    # reservoir Gram matrices against the SVD solution: W_out error ~ 3e-14 / ratio (ratio 2.8e-6 and error
    # 6e-9 for a cfg3 pilot with the default noise, 3.7e-9 / 2.4e-6 with noise = 0, 9.6e-10 / 1.6e-5, 2.7e-12 / 1e-2)
    PIVOT_RATIO_MIN = 1e-9

    def train_readout(self, ext, teachers, transient=0, shared=False, stable_fallback=False):
        """fp64 normal equations with lambda = 0 + Cholesky, reproducing the
        reference's pinv solution (libs/pyESN.py:191-192).  One
        readout per frame, or ONE readout over all frames when `shared`.
        Returns (W_out [G, n_out, P] fp64, info [G] int32).

        The reference solves with an SVD (`np.linalg.pinv`, rcond 1e-15), which survives extended states
        with cond(E) ~ 1e9 (noise = 0, long frames) where the normal equations silently lose every digit.
        The Cholesky kernel reports its pivot range (`self.last_pivot_ratio`, [G]); with
        `stable_fallback` the problems whose factorisation failed (info != 0) or whose pivot ratio is
        below PIVOT_RATIO_MIN are re-solved on the device by an fp64 SVD pseudo-inverse of E with the
        reference's rcond (one host sync; the drop-in ESN.fit uses it, the batched hot path does not)."""
        B, T, P = ext.shape
        m = T - int(transient)
        dual = (m < P) and not shared
        G_n = m if dual else P
        nprob = 1 if shared else B
        teachers = teachers.to(self.device).contiguous()
        if teachers.dim() == 2:
            teachers = teachers.unsqueeze(0)
        aff = self._aff[ESN_F64]
        G = torch.empty((nprob, G_n, G_n), dtype=torch.float64, device=self.device)
        rhs = torch.empty((nprob, G_n, self.n_out), dtype=torch.float64, device=self.device)
        check(self.lib.esn_gram_f64(ptr(ext), _CODE[ext.dtype], ptr(teachers), _CODE[teachers.dtype],
                                    ptr(aff["t_scale"]), ptr(aff["t_shift"]), B, T, P, self.n_out,
                                    int(transient), int(dual), int(shared), 0, ptr(G), ptr(rhs), _stream()),
              "esn_gram_f64")
        W_out, info = self.solve_readout(G, rhs, ext if dual else None, transient)
        if stable_fallback:
            bad = (info != 0) | ~(self.last_pivot_ratio >= self.PIVOT_RATIO_MIN)
            if bool(bad.any()):
                W_out, info = self._pinv_fallback(ext, teachers, transient, shared, bad, W_out, info)
        return W_out, info

    def _pinv_fallback(self, ext, teachers, transient, shared, bad, W_out, info):
        """Ill-conditioned problems only: W_out = (pinv(E[transient:]) Y)^T in fp64 with rcond = 1e-15, the
        reference's own formula (libs/pyESN.py:191-192), via the device SVD."""
        aff = self._aff[ESN_F64]
        E = ext[:, transient:, :].to(torch.float64)
        Y = teachers[:, transient:, :].to(torch.float64) * aff["t_scale"] + aff["t_shift"]
        if shared:
            E, Y = E.reshape(1, -1, E.shape[-1]), Y.reshape(1, -1, Y.shape[-1])
        W_out, info = W_out.clone(), info.clone()
        for g in torch.nonzero(bad).flatten().tolist():
            W_out[g] = (torch.linalg.pinv(E[g], rtol=1e-15) @ Y[g]).T
            info[g] = 0
        return W_out, info

    def gram(self, ext, teachers, transient=0):
        """Shared-readout partial sums (G [P,P], R [P,n_out]) of this rank's
        frames; allreduce them over ranks, then `solve_readout`."""
        B, T, P = ext.shape
        teachers = teachers.to(self.device).contiguous()
        aff = self._aff[ESN_F64]
        G = torch.empty((1, P, P), dtype=torch.float64, device=self.device)
        rhs = torch.empty((1, P, self.n_out), dtype=torch.float64, device=self.device)
        check(self.lib.esn_gram_f64(ptr(ext), _CODE[ext.dtype], ptr(teachers), _CODE[teachers.dtype],
                                    ptr(aff["t_scale"]), ptr(aff["t_shift"]), B, T, P, self.n_out,
                                    int(transient), 0, 1, 0, ptr(G), ptr(rhs), _stream()), "esn_gram_f64")
        return G, rhs

    def gram_flat(self, ext, teachers, transient=0):
        """As `gram`, in ONE flat fp64 buffer [P*P + P*n_out] (what goes on the wire when the frames of a shared
        readout are spread over ranks: dist.GramReducer).  Returns (flat, G view [1,P,P], R view [1,P,n_out])."""
        B, T, P = ext.shape
        teachers = teachers.to(self.device).contiguous()
        aff = self._aff[ESN_F64]
        flat = torch.empty((P * P + P * self.n_out,), dtype=torch.float64, device=self.device)
        G, rhs = flat[:P * P].view(1, P, P), flat[P * P:].view(1, P, self.n_out)
        check(self.lib.esn_gram_f64(ptr(ext), _CODE[ext.dtype], ptr(teachers), _CODE[teachers.dtype],
                                    ptr(aff["t_scale"]), ptr(aff["t_shift"]), B, T, P, self.n_out,
                                    int(transient), 0, 1, 0, ptr(G), ptr(rhs), _stream()), "esn_gram_f64")
        return flat, G, rhs

    def train_shared_readout(self, inputs, teachers, transient=0, precision="fp64", chunks=4, seed=0,
                             noise_uniforms=None, su_exp=None, y_absmax=None):
        """ONE readout from this rank's pilots AND those of every other rank (BASELINE.json configs[4]: "readout Gram
        allreduced over NVLink"): chunked harvest -> partial normal equations -> asynchronous all-reduce per chunk
        (hidden behind the next chunk's harvest) -> the same Cholesky on every rank.  Returns (W_out [1,n_out,P],
        info, bytes all-reduced)."""
        from . import dist as D
        B = inputs.shape[0]
        red = D.GramReducer()
        P = self.P
        for k in range(chunks):
            b0, b1 = D.shard_range(B, k, chunks)
            if b1 <= b0:
                continue
            nu = None if noise_uniforms is None else noise_uniforms[b0:b1]
            ext = self.harvest(inputs[b0:b1], teachers[b0:b1], precision=precision, noise_uniforms=nu, seed=seed + k,
                               su_exp=su_exp, y_absmax=y_absmax)
            flat, _, _ = self.gram_flat(ext, teachers[b0:b1], transient)
            red.add(flat)
            del ext
        flat = red.finish()
        W_out, info = self.solve_readout(flat[:P * P].view(1, P, P), flat[P * P:].view(1, P, self.n_out))
        return W_out, info, red.bytes

    def solve_readout(self, G, rhs, ext_for_dual=None, transient=0):
        nprob, n, _ = G.shape
        info = torch.zeros((nprob,), dtype=torch.int32, device=self.device)
        G = G.contiguous()
        rhs = rhs.contiguous()
        piv = torch.empty((nprob, 2), dtype=torch.float64, device=self.device)
        check(self.lib.esn_cholesky_solve_piv_f64(ptr(G), ptr(rhs), nprob, n, self.n_out, ptr(info), ptr(piv),
                                                  _stream()), "esn_cholesky_solve_piv_f64")
        self.last_pivot_ratio = piv[:, 0] / piv[:, 1]
        W_out = torch.empty((nprob, self.n_out, self.P), dtype=torch.float64, device=self.device)
        if ext_for_dual is not None:
            B, T, P = ext_for_dual.shape
            check(self.lib.esn_readout_from_dual_f64(ptr(ext_for_dual), _CODE[ext_for_dual.dtype], ptr(rhs),
                                                     B, T, P, self.n_out, int(transient), ptr(W_out), _stream()),
                  "esn_readout_from_dual_f64")
        else:
            check(self.lib.esn_transpose_rhs_f64(ptr(rhs), nprob, self.P, self.n_out, ptr(W_out), _stream()),
                  "esn_transpose_rhs_f64")
        return W_out, info

    def apply_readout(self, ext, W_out, group_ids=None):
        """Train-set prediction (libs/pyESN.py:212-213) on all T rows."""
        B, T, P = ext.shape
        code = _CODE[ext.dtype]
        W_out = W_out.to(device=self.device, dtype=ext.dtype).contiguous()
        if W_out.dim() == 2:
            W_out = W_out.unsqueeze(0)
        if group_ids is None and W_out.shape[0] == B and B > 1:
            group_ids = torch.arange(B, dtype=torch.int32, device=self.device)
        if group_ids is not None:
            group_ids = group_ids.to(device=self.device, dtype=torch.int32).contiguous()
        aff = self._aff[code]
        pred = torch.empty((B, T, self.n_out), dtype=ext.dtype, device=self.device)
        check(self.lib.esn_apply_readout(code, ptr(ext), ptr(W_out), ptr(group_ids), ptr(aff["t_scale"]),
                                         ptr(aff["t_shift"]), B, T, P, self.n_out, ptr(pred), _stream()),
              "esn_apply_readout")
        return pred
