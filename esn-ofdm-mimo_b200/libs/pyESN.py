"""Drop-in `pyESN` module backed by the B200 engine.

Same public surface as the reference module (`/root/reference/libs/pyESN.py`):
`ESN(...)` with the identical constructor signature, `.fit()`, `.predict()`,
plus the module-level `correct_dimensions` and `identity`.  Put this directory
on PYTHONPATH instead of the reference's `libs/` and the System Model 2 demos
run unchanged.

What stays on the host, and why: reservoir initialisation draws from numpy's
MT19937 and rescales by `eigvals` (reference :93-109) -- it is setup, runs once
per ESN, and must be bit-identical for seed parity.  The state-noise uniforms
are drawn from the same `random_state_` in the reference's order (T-1 rows per
`fit`, T rows per `predict`, reference :125) and shipped to the GPU, so results
track the reference on identical seeds.  Everything else -- the time loops, the
readout solve, the train-set prediction -- runs in CUDA (no CPU fallback).

Additive API (not in the reference): `fit_batched` / `predict_batched` on torch
CUDA tensors, `precision=` ('fp64' default for the single-frame calls and for
training; 'auto' for batched detection: tensor cores at large batch, cluster /
SIMT fp32 kernels at small batch).
"""
import os
import sys

import numpy as np

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)


def correct_dimensions(s, targetlength):
    """None -> None; scalar -> vector of `targetlength` copies; 1-D of the right
    length -> itself; anything else -> ValueError (reference :4-24)."""
    if s is None:
        return None
    s = np.array(s)
    if s.ndim > 1:
        raise ValueError("Invalid argument")
    if s.ndim == 1:
        if len(s) != targetlength:
            raise ValueError("arg must have length " + str(targetlength))
        return s
    return np.array([s] * targetlength)


def identity(x):
    return x


def _as_2d(a):
    a = np.asarray(a)
    return a.reshape(len(a), -1) if a.ndim < 2 else a


_RADIUS_CACHE = {}                        # digest of the unscaled matrix -> max |eigenvalue| (last 16 matrices)


def _spectral_radius(W):
    """`np.max(np.abs(np.linalg.eigvals(W)))` exactly as the reference computes it (libs/pyESN.py:99: host LAPACK, so
    the scaled weights stay bit-identical), remembered per matrix: scripts that rebuild an ESN from the same seed --
    a sweep over spectral radii, a demo with `random_state=42` per coherence block -- draw the same unscaled matrix
    again, and `eigvals` (0.4 s at 512 neurons, 3.5 s at 2048) is all that construction costs."""
    import hashlib
    key = (W.shape, hashlib.blake2b(np.ascontiguousarray(W).view(np.uint8), digest_size=16).digest())
    r = _RADIUS_CACHE.get(key)
    if r is None:
        r = float(np.max(np.abs(np.linalg.eigvals(W))))
        if len(_RADIUS_CACHE) >= 16:
            _RADIUS_CACHE.pop(next(iter(_RADIUS_CACHE)))
        _RADIUS_CACHE[key] = r
    return r


class ESN():
    """Echo State Network with the reference's constructor (reference :33-91)."""

    def __init__(self, n_inputs, n_outputs, n_reservoir=200,
                 spectral_radius=0.95, sparsity=0, noise=0.001, input_shift=None,
                 input_scaling=None, teacher_forcing=True, feedback_scaling=None,
                 teacher_scaling=None, teacher_shift=None,
                 out_activation=identity, inverse_out_activation=identity,
                 random_state=None, silent=True, precision="fp64"):
        if out_activation is not identity or inverse_out_activation is not identity:
            # no reference call site uses anything else; refuse rather than diverge silently
            raise NotImplementedError("esn_b200 supports the identity output activation only")
        self.n_inputs, self.n_outputs, self.n_reservoir = n_inputs, n_outputs, n_reservoir
        self.spectral_radius, self.sparsity, self.noise = spectral_radius, sparsity, noise
        self.input_shift = correct_dimensions(input_shift, n_inputs)
        self.input_scaling = correct_dimensions(input_scaling, n_inputs)
        self.teacher_scaling, self.teacher_shift = teacher_scaling, teacher_shift
        self.out_activation, self.inverse_out_activation = out_activation, inverse_out_activation
        self.random_state = random_state
        self.teacher_forcing, self.silent = teacher_forcing, silent
        self.precision = precision
        # feedback_scaling is accepted and ignored, exactly as the reference does (:35)
        if isinstance(random_state, np.random.RandomState):
            self.random_state_ = random_state
        elif random_state:
            try:
                self.random_state_ = np.random.RandomState(random_state)
            except TypeError as e:
                raise Exception("Invalid seed: " + str(e))
        else:                               # None, 0, ... -> numpy's global generator (:86-87)
            self.random_state_ = np.random.mtrand._rand
        self._dev = None
        self.initweights()

    # ---- host-side setup (bit-identical weights) -------------------------
    def initweights(self):
        """Draw order is part of the contract (reference :93-109)."""
        rs, N = self.random_state_, self.n_reservoir
        W = rs.rand(N, N) - 0.5
        W[rs.rand(N, N) < self.sparsity] = 0
        W *= self.spectral_radius / _spectral_radius(W)
        self.W = W
        self.W_in = rs.rand(N, self.n_inputs) * 2 - 1
        self.W_feedb = rs.rand(N, self.n_outputs) * 2 - 1
        self._dev = None

    def _engine(self):
        """The device-side reservoir.  The reference reads W, W_in, W_feedb, the scalings, `noise` and
        `teacher_forcing` live on every call (reference :111-152), so attributes changed after construction
        must take effect: the weights are re-uploaded when they were replaced or edited, the affine maps and
        the noise amplitude are refreshed in place (no upload)."""
        def vec(v):
            return None if v is None else np.asarray(v, dtype=np.float64).tobytes()
        wsig = (id(self.W), id(self.W_in), id(self.W_feedb), self.W.shape, float(self.W.sum()),
                float(self.W_in.sum()), float(self.W_feedb.sum()), bool(self.teacher_forcing))
        asig = (vec(self.input_scaling), vec(self.input_shift), vec(self.teacher_scaling),
                vec(self.teacher_shift), float(self.noise))
        if self._dev is None or self._dev_wsig != wsig:
            from esn_b200 import Reservoir
            self._dev = Reservoir(self.W, self.W_in, self.W_feedb,
                                  input_scaling=self.input_scaling, input_shift=self.input_shift,
                                  teacher_scaling=self.teacher_scaling, teacher_shift=self.teacher_shift,
                                  noise=self.noise, teacher_forcing=self.teacher_forcing)
            self._dev_wsig, self._dev_asig = wsig, asig
        elif self._dev_asig != asig:
            self._dev = self._dev.with_affine(self.input_scaling, self.input_shift, self.teacher_scaling,
                                              self.teacher_shift, self.noise)
            self._dev_asig = asig
        return self._dev

    # ---- the affine maps, kept for API parity (reference :127-152) -------
    def _scale_inputs(self, inputs):
        if self.input_scaling is not None:
            inputs = inputs * self.input_scaling
        if self.input_shift is not None:
            inputs = inputs + self.input_shift
        return inputs

    def _scale_teacher(self, teacher):
        if self.teacher_scaling is not None:
            teacher = teacher * self.teacher_scaling
        if self.teacher_shift is not None:
            teacher = teacher + self.teacher_shift
        return teacher

    def _unscale_teacher(self, teacher_scaled):
        if self.teacher_shift is not None:
            teacher_scaled = teacher_scaled - self.teacher_shift
        if self.teacher_scaling is not None:
            teacher_scaled = teacher_scaled / self.teacher_scaling
        return teacher_scaled

    # ---- fit / predict ----------------------------------------------------
    def _device_rng(self):
        """The noise generator continued on the device (esn_b200.noise.DeviceRandomState): the same doubles as
        `random_state_.rand(...)`, without the host draw.  None for a generator that is not numpy's MT19937
        RandomState (then the rows are drawn on the host, as the reference does)."""
        import torch
        from esn_b200.noise import DeviceRandomState
        try:
            return DeviceRandomState(self.random_state_), (torch.float64 if self.precision == "fp64" else torch.float32)
        except (ValueError, AttributeError, TypeError):
            return None, None

    def fit(self, inputs, outputs, transient=0, inspect=False):
        """Harvest states under teacher forcing, solve the readout, return the
        train-set prediction on all rows (reference :154-216)."""
        import torch
        inputs, outputs = _as_2d(inputs), _as_2d(outputs)
        T = inputs.shape[0]
        eng = self._engine()
        if not self.silent:
            print("harvesting states...")
        # one rand(N) row per step n = 1..T-1, drawn even when noise == 0 so the
        # generator stays in step with the reference
        drng, dt = self._device_rng()
        if drng is not None:
            try:
                uni = drng.rand(max(T - 1, 0), self.n_reservoir, dt)
                ext = eng.harvest(inputs[None], outputs[None], precision=self.precision,
                                  noise_uniforms=uni[None] if self.noise != 0 and T > 1 else None)
            finally:
                drng.finalize()
        else:
            uni = self.random_state_.rand(max(T - 1, 0), self.n_reservoir)
            ext = eng.harvest(inputs[None], outputs[None], precision=self.precision,
                              noise_uniforms=uni[None] if self.noise != 0 and T > 1 else None)
        if not self.silent:
            print("fitting...")
        teach = torch.from_numpy(np.ascontiguousarray(outputs, dtype=np.float64)[None])
        W_out, info = eng.train_readout(ext, teach, transient, stable_fallback=True)
        if int(info[0]) != 0:
            raise np.linalg.LinAlgError(
                f"readout Gram matrix is not positive definite (pivot {int(info[0])}): "
                "the extended states are rank deficient")
        self.W_out = W_out[0].cpu().numpy()
        self.laststate = ext[0, -1, :self.n_reservoir].to(torch.float64).cpu().numpy()
        self.lastinput = inputs[-1, :]
        self.lastoutput = self._scale_teacher(outputs)[-1, :]
        if inspect:
            from matplotlib import pyplot as plt
            E = ext[0].cpu().numpy()
            plt.figure(figsize=(E.shape[0] * 0.0025, E.shape[1] * 0.01))
            plt.imshow(E.T, aspect='auto', interpolation='nearest')
            plt.colorbar()
        if not self.silent:
            print("training error:")
        pred_train = eng.apply_readout(ext, W_out.to(ext.dtype))[0].to(torch.float64).cpu().numpy()
        if not self.silent:
            print(np.sqrt(np.mean((pred_train - outputs) ** 2)))
        return pred_train

    def predict(self, inputs, transient=0, continuation=True):
        """Free-running prediction with output feedback (reference :218-255);
        does not modify the ESN."""
        inputs = _as_2d(inputs)
        T = inputs.shape[0]
        x0 = y0 = None
        if continuation:
            x0, y0 = self.laststate[None], np.asarray(self.lastoutput, dtype=np.float64).reshape(1, -1)
        W_out = self.W_out                      # AttributeError before fit(), as the reference
        eng = self._engine()
        import torch
        drng, dt = self._device_rng()
        if drng is not None:
            try:
                uni = drng.rand(T, self.n_reservoir, dt)
                y = eng.predict(inputs[None], W_out[None], transient=transient, x0=x0, y0=y0,
                                precision=self.precision,
                                noise_uniforms=uni[None] if self.noise != 0 else None)
                y = y[0].to(torch.float64).cpu().numpy()
            finally:
                drng.finalize()
            return y
        uni = self.random_state_.rand(T, self.n_reservoir)
        y = eng.predict(inputs[None], W_out[None], transient=transient, x0=x0, y0=y0,
                        precision=self.precision,
                        noise_uniforms=uni[None] if self.noise != 0 else None)
        return y[0].to(torch.float64).cpu().numpy()

    def fit_predict_many(self, cases):
        """Batched form of the trainers' delay scan (reference
        libs/helper_mimo_esn_generic.py:67-81, libs/HelpFunc.py:109-154): for every
        (inputs, outputs, transient) in `cases`, in order, the reference calls
        fit(inputs, outputs, transient) and then predict(inputs, transient,
        continuation=False).  Here all candidates go through ONE batched harvest, ONE
        batched Gram + Cholesky and ONE batched predict launch.  The state noise is drawn
        from `random_state_` in exactly the reference's order (fit rows, then predict rows,
        candidate after candidate), so results equal the sequential calls.  Returns the list
        of predictions; W_out / laststate / lastoutput end up as after the last fit."""
        import torch
        cases = [(_as_2d(i), _as_2d(o), int(t)) for i, o, t in cases]
        n, N = len(cases), self.n_reservoir
        Ts = [c[0].shape[0] for c in cases]
        ms = {T - c[2] for T, c in zip(Ts, cases)}
        if n == 0:
            return []
        if len(ms) != 1 or min(Ts) < 2:                # ragged row windows: plain sequential calls
            out = []
            for i, o, t in cases:
                self.fit(i, o, t)
                out.append(self.predict(i, t, continuation=False))
            return out
        m, Tp = ms.pop(), max(Ts)
        X_in = np.zeros((n, Tp, self.n_inputs))
        X_out = np.zeros((n, Tp, self.n_outputs))
        eng = self._engine()
        drng, dt = self._device_rng()
        if drng is not None:
            try:
                uni_fit = torch.full((n, Tp - 1, N), 0.5, dtype=dt, device=drng.device)   # 0.5 = no noise on the padded tail
                uni_pred = torch.full((n, Tp, N), 0.5, dtype=dt, device=drng.device)
                for k, (i, o, t) in enumerate(cases):
                    T = Ts[k]
                    X_in[k, :T], X_out[k, :T] = i, o
                    uni_fit[k, :T - 1] = drng.rand(T - 1, N, dt)
                    uni_pred[k, :T] = drng.rand(T, N, dt)
            finally:
                drng.finalize()
        else:
            uni_fit = np.full((n, Tp - 1, N), 0.5)
            uni_pred = np.full((n, Tp, N), 0.5)
            for k, (i, o, t) in enumerate(cases):
                T = Ts[k]
                X_in[k, :T], X_out[k, :T] = i, o
                uni_fit[k, :T - 1] = self.random_state_.rand(T - 1, N)
                uni_pred[k, :T] = self.random_state_.rand(T, N)
        noisy = self.noise != 0
        ext = eng.harvest(X_in, X_out, precision=self.precision, noise_uniforms=uni_fit if noisy else None)
        dev = ext.device
        rows = torch.tensor([[c[2] + r for r in range(m)] for c in cases], device=dev)       # [n, m]
        sel = torch.arange(n, device=dev)[:, None]
        teach = torch.from_numpy(X_out).to(dev)
        W_out, info = eng.train_readout(ext[sel, rows].contiguous(), teach[sel, rows].contiguous(), 0,
                                        stable_fallback=True)
        if int(info.abs().max()) != 0:
            bad = int(torch.nonzero(info)[0])
            raise np.linalg.LinAlgError(
                f"readout Gram matrix of candidate {bad} is not positive definite (pivot {int(info[bad])})")
        y = eng.predict(X_in, W_out, transient=0, group_ids=torch.arange(n, dtype=torch.int32, device=dev),
                        precision=self.precision, noise_uniforms=uni_pred if noisy else None)
        y = y.to(torch.float64).cpu().numpy()
        last = n - 1
        self.W_out = W_out[last].cpu().numpy()
        self.laststate = ext[last, Ts[last] - 1, :N].to(torch.float64).cpu().numpy()
        self.lastinput = cases[last][0][-1, :]
        self.lastoutput = self._scale_teacher(cases[last][1])[-1, :]
        return [y[k, cases[k][2]:Ts[k]] for k in range(n)]

    # ---- additive batched API (torch CUDA tensors) ------------------------
    def fit_batched(self, inputs, outputs, transient=0, precision="fp64", noise_uniforms=None,
                    seed=0, shared=False):
        """Train one readout per frame (or one shared readout) for B pilot
        frames [B,T,n_in]/[B,T,n_out].  Returns (W_out [G,n_out,P] fp64 device
        tensor, info [G]).  State noise: `noise_uniforms` [B,T-1,N] if given,
        else the device counter stream `seed`."""
        eng = self._engine()
        ext = eng.harvest(inputs, outputs, precision=precision, noise_uniforms=noise_uniforms, seed=seed)
        return eng.train_readout(ext, outputs, transient, shared=shared)

    def predict_batched(self, inputs, W_out, transient=0, group_ids=None, precision="auto",
                        noise_uniforms=None, seed=0, x0=None, y0=None):
        """Detect B frames [B,T,n_in] in one launch; frame b uses readout
        W_out[group_ids[b]].  Returns y [B,T-transient,n_out] on the device.  precision='auto' picks the fastest
        kernel for the batch size, reservoir size and readout layout (tensor cores at large batch, the
        cluster / SIMT fp32 kernels at small batch; esn_b200.engine.Reservoir.auto_predict_path)."""
        return self._engine().predict(inputs, W_out, transient=transient, group_ids=group_ids,
                                      precision=precision, noise_uniforms=noise_uniforms, seed=seed,
                                      x0=x0, y0=y0)
