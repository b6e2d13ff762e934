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
