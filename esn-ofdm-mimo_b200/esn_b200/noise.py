"""Host restatement (numpy, integer-exact) of the device counter hash that
draws the reservoir's state noise when no host noise tensor is supplied
(csrc/common.cuh: esn_mix32 / esn_noise_key / esn_noise_uniform).  Used by the
tests to hand the oracle the very stream the kernel used."""
import numpy as np

_M = np.uint64(0xFFFFFFFF)


def _mix32(x):
    x = x & _M
    x ^= x >> np.uint64(16)
    x = (x * np.uint64(0x7FEB352D)) & _M
    x ^= x >> np.uint64(15)
    x = (x * np.uint64(0x846CA68B)) & _M
    x ^= x >> np.uint64(16)
    return x


def device_noise_uniforms(seed, B, steps, N, first_frame=0):
    """uniforms[b, row, neuron] in [0,1), exactly as the kernels draw them: one key
    per (frame, row), one hash per neuron pair; even neuron = low 16 bits, odd = high 16."""
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    lo, hi = np.uint64(seed & 0xFFFFFFFF), np.uint64(seed >> 32)
    b = (np.arange(B, dtype=np.uint64) + np.uint64(first_frame))[:, None, None]
    r = np.arange(steps, dtype=np.uint64)[None, :, None]
    n = np.arange(N, dtype=np.uint64)[None, None, :]
    odd = (n & np.uint64(1)).astype(bool)
    k = _mix32(lo + ((np.uint64(0x9E3779B9) * b) & _M))
    k = _mix32(k ^ ((((r * np.uint64(0x85EBCA6B)) & _M) + hi) & _M))
    x = (k + (((n >> np.uint64(1)) * np.uint64(0xC2B2AE35)) & _M)) & _M
    x ^= x >> np.uint64(16)
    m = x * np.uint64(0x9E3779B1)                      # 32 x 32 -> 64 bit product, high word folded onto the low
    h = (m & _M) ^ (m >> np.uint64(32))
    bits = np.where(odd, h >> np.uint64(16), h & np.uint64(0xFFFF))
    return bits.astype(np.float64) / 65536.0


class DeviceRandomState:
    """The MT19937 stream of a numpy `RandomState` continued on the device (esn_mt19937_uniforms): `rand(rows,
    cols)` returns the very doubles `rs.rand(rows, cols)` would, as a CUDA tensor, without the 1.3 ms host draw
    and the 2 MB upload a cfg3 frame costs; `finalize()` reads the advanced state back (one small copy) and
    installs it in `rs`, so the caller's generator -- possibly numpy's global one -- continues as after the
    reference's own draws.  Several draws chain on the device between upload and finalize."""

    def __init__(self, rs, device="cuda"):
        import torch
        from . import _lib
        st = rs.get_state(legacy=True)
        if st[0] != "MT19937":
            raise ValueError("not an MT19937 RandomState")
        self._rs, self._tail = rs, (st[3], st[4])
        self._lib, self._torch = _lib, torch
        self.device = torch.device(device)
        host = np.empty(625, dtype=np.uint32)
        host[:624] = st[1]
        host[624] = st[2]
        self._state = torch.from_numpy(host.view(np.int32)).to(self.device)
        self._keep = []

    def rand(self, rows, cols, dtype=None):
        torch, L = self._torch, self._lib
        dtype = dtype or torch.float64
        count = int(rows) * int(cols)
        out = torch.empty((int(rows), int(cols)), dtype=dtype, device=self.device)
        if count == 0:
            return out
        words = torch.empty((2 * count,), dtype=torch.int32, device=self.device)
        code = L.ESN_F64 if dtype == torch.float64 else L.ESN_F32
        L.check(L.load().esn_mt19937_uniforms(L.ptr(self._state), count, code, L.ptr(words), L.ptr(out),
                                             torch.cuda.current_stream().cuda_stream), "esn_mt19937_uniforms")
        self._keep.append(words)
        return out

    def finalize(self):
        h = self._state.cpu().numpy().view(np.uint32)
        self._rs.set_state(("MT19937", h[:624].copy(), int(h[624]), self._tail[0], self._tail[1]))
        self._keep.clear()
