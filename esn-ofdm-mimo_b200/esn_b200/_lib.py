"""ctypes binding of the C-ABI shared library (include/esn_b200.h).

The library is built in-tree by `__graft_entry__.build()` (nvcc, sm_100a) as
`esn-ofdm-mimo_b200/libesn_b200.so`.  There is no CPU fallback: if the library
is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ESN_B200_LIB") or os.path.join(os.path.dirname(_HERE), "libesn_b200.so")   # override: kernel experiments

ESN_F32, ESN_F64 = 0, 1
MODE_HARVEST, MODE_PREDICT = 0, 1
ESN_MAX_OUT, ESN_MAX_IN = 16, 64

_ERRORS = {-1: "ESN_E_BADARG (bad argument)", -2: "ESN_E_TOOLARGE (shape does not fit one SM)",
           -3: "ESN_E_NODEVICE (no CUDA device)", -4: "ESN_E_UNSUPPORTED (shape not supported by this path)"}


class EsnB200Error(RuntimeError):
    pass


class RecurrenceArgs(C.Structure):
    _fields_ = [
        ("dtype", C.c_int32), ("mode", C.c_int32), ("B", C.c_int32), ("T", C.c_int32),
        ("N", C.c_int32), ("n_in", C.c_int32), ("n_out", C.c_int32),
        ("N_pad", C.c_int32), ("K_aug_pad", C.c_int32), ("transient", C.c_int32),
        ("feedback", C.c_int32), ("n_groups", C.c_int32),
        ("noise_amp", C.c_double), ("seed", C.c_uint64),
        ("Wt_aug", C.c_void_p), ("inp", C.c_void_p), ("in_scale", C.c_void_p),
        ("in_shift", C.c_void_p), ("teacher", C.c_void_p), ("t_scale", C.c_void_p),
        ("t_shift", C.c_void_p), ("W_out", C.c_void_p), ("group_ids", C.c_void_p),
        ("x0", C.c_void_p), ("y0", C.c_void_p), ("noise_uniforms", C.c_void_p),
        ("ext_out", C.c_void_p), ("y_out", C.c_void_p), ("workspace", C.c_void_p),
    ]


class TcPredictArgs(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("T", C.c_int32), ("N", C.c_int32), ("n_in", C.c_int32), ("n_out", C.c_int32),
        ("transient", C.c_int32), ("feedback", C.c_int32), ("su_exp", C.c_int32), ("sy_exp", C.c_int32),
        ("n_groups", C.c_int32),
        ("noise_amp", C.c_double), ("seed", C.c_uint64),
        ("weights", C.c_void_p), ("readouts", C.c_void_p), ("yscale", C.c_void_p), ("inp", C.c_void_p),
        ("in_scale", C.c_void_p), ("in_shift", C.c_void_p), ("t_scale", C.c_void_p), ("t_shift", C.c_void_p),
        ("group_ids", C.c_void_p), ("x0", C.c_void_p), ("y0", C.c_void_p), ("noise_uniforms", C.c_void_p),
        ("ext_out", C.c_void_p), ("y_out", C.c_void_p), ("timeline", C.c_void_p),
        ("reserved", C.c_int32), ("teacher", C.c_void_p),
    ]


class TcsArgs(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("T", C.c_int32), ("N", C.c_int32), ("n_in", C.c_int32), ("n_out", C.c_int32),
        ("transient", C.c_int32), ("feedback", C.c_int32), ("su_exp", C.c_int32), ("sy_exp", C.c_int32),
        ("n_groups", C.c_int32), ("accumulators", C.c_int32), ("ring_a", C.c_int32), ("ring_b", C.c_int32),
        ("noise_amp", C.c_double), ("seed", C.c_uint64),
        ("weights", C.c_void_p), ("wo_x", C.c_void_p), ("wo_u", C.c_void_p), ("inp", C.c_void_p),
        ("in_scale", C.c_void_p), ("in_shift", C.c_void_p), ("t_scale", C.c_void_p), ("t_shift", C.c_void_p),
        ("group_ids", C.c_void_p), ("x0", C.c_void_p), ("y0", C.c_void_p), ("noise_uniforms", C.c_void_p),
        ("ext_out", C.c_void_p), ("y_out", C.c_void_p), ("teacher", C.c_void_p), ("workspace", C.c_void_p),
        ("timeline", C.c_void_p),
    ]


# name -> (restype, argtypes); every symbol include/esn_b200.h declares
_vp, _i, _d = C.c_void_p, C.c_int, C.c_double
SIGNATURES = {
    "esn_version": (_i, []),
    "esn_device_info": (_i, [C.c_char_p, _i, C.POINTER(_i), C.POINTER(_i)]),
    "esn_noise_uniform_host": (C.c_float, [C.c_uint64, C.c_uint, C.c_uint, C.c_uint]),
    "esn_mt19937_uniforms": (_i, [_vp, C.c_longlong, _i, _vp, _vp, _vp]),
    "esn_pad_sizes": (_i, [_i, _i, _i, C.POINTER(_i), C.POINTER(_i)]),
    "esn_set_small_batch_limit": (_i, [_i]),
    "esn_recurrence_run": (_i, [C.POINTER(RecurrenceArgs), _vp]),
    "esn_tc_supported": (_i, [_i, _i, _i]),
    "esn_tc_weight_bytes": (C.c_longlong, [_i, _i]),
    "esn_tc_readout_bytes": (C.c_longlong, [_i, _i]),
    "esn_tc_prepare_weights": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "esn_tc_prepare_readout": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "esn_tc_predict": (_i, [C.POINTER(TcPredictArgs), _vp]),
    "esn_tc_set_acc_k0": (_d, [_d]),
    "esn_tc_acc_k0": (_d, []),
    "esn_tcs_supported": (_i, [_i, _i, _i]),
    "esn_tcs_workspace_bytes": (C.c_longlong, [_i, _i]),
    "esn_tcs_readout_floats": (C.c_longlong, [_i, _i, C.POINTER(C.c_longlong)]),
    "esn_tcs_prepare_readout": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "esn_tcs_run": (_i, [C.POINTER(TcsArgs), _vp]),
    "esn_tcr_supported": (_i, [_i, _i, _i]),
    "esn_tcr_run": (_i, [C.POINTER(TcsArgs), _vp]),
    "esn_gram_f64": (_i, [_vp, _i, _vp, _i, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "esn_cholesky_solve_f64": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp]),
    "esn_cholesky_solve_piv_f64": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp]),
    "esn_readout_from_dual_f64": (_i, [_vp, _i, _vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "esn_transpose_rhs_f64": (_i, [_vp, _i, _i, _i, _vp, _vp]),
    "esn_apply_readout": (_i, [_i, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "ofdm_unpack_fft_demap": (_i, [_i, _vp, _i, _i, _i, _i, _vp, _i, _i, _vp, _vp, _vp, _d, _vp, _vp]),
    "ofdm_rx_fft": (_i, [_i, _vp, _i, _i, _i, _i, _vp, _vp]),
    "ofdm_equalize": (_i, [_i, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _i, _vp, _i, _vp, _vp]),
    "ofdm_chanest": (_i, [_i, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _i, _d, _vp, _vp, _vp]),
    "ofdm_synth_frames": (_i, [_i, _vp, _vp, _vp, _vp, _vp, _vp, _d, C.c_uint64, _i, _i, _i, _i, _i, _i, _i, _i,
                               _vp, _vp, _vp, _vp]),
    "ofdm_demap_count": (_i, [_i, _vp, _i, _i, _i, _i, _vp, _vp, _d, _vp, _vp]),
    "ofdm_soft_demap": (_i, [_i, _vp, _i, _i, _i, _i, _vp, _vp, _d, _vp, _vp, _vp]),
    "ofdm_llr_calibrate": (_i, [_i, _vp, _vp, _i, _i, _i, _i, _i, _d, _d, _vp, _vp]),
}

_lib = None


def load():
    """Load libesn_b200.so (once) and type its entry points."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EsnB200Error(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def check(rc, what):
    if rc == 0:
        return
    if rc < 0:
        raise EsnB200Error(f"{what}: {_ERRORS.get(rc, rc)}")
    raise EsnB200Error(f"{what}: CUDA error {rc}")


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else C.c_void_p(t.data_ptr())
