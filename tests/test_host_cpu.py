"""CPU-only checks: the C-ABI library loads and exports every symbol the header
declares, the drop-in modules keep the reference's host-side behaviour
(constructor checks, RNG consumption, error types), the noise-hash restatement
matches the library, and the multi-rank host logic works under gloo."""
import os
import re
import sys

import numpy as np
import pytest

import cases
from conftest import ROOT


def _build():
    import __graft_entry__ as g
    return g.build()


def test_library_exports_every_declared_symbol():
    lib_path = _build()
    import ctypes
    lib = ctypes.CDLL(lib_path)
    hdr = open(os.path.join(ROOT, "include", "esn_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = re.findall(r"^\s*(?:int|float|double|long long)\s+((?:esn|ofdm)_\w+)\s*\(", hdr, flags=re.M)
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in esn_b200.h but not exported"
    from esn_b200 import _lib
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)
    assert lib.esn_version() >= 1


def test_pad_sizes_and_argument_errors():
    _build()
    import ctypes as C
    from esn_b200 import _lib
    lib = _lib.load()
    a, b = C.c_int(), C.c_int()
    assert lib.esn_pad_sizes(512, 16, 8, C.byref(a), C.byref(b)) == 0
    assert (a.value, b.value) == (512, 544)
    assert lib.esn_pad_sizes(100, 4, 4, C.byref(a), C.byref(b)) == 0
    assert (a.value, b.value) == (128, 112)
    assert lib.esn_pad_sizes(0, 4, 4, C.byref(a), C.byref(b)) == -1
    args = _lib.RecurrenceArgs()           # all zero -> bad argument, no launch
    assert lib.esn_recurrence_run(C.byref(args), None) == -1
    assert lib.esn_cholesky_solve_f64(None, None, 1, 4, 1, None, None) == -1
    assert lib.ofdm_rx_fft(0, None, 1, 64, 7, 2, None, None) == -1


def test_noise_hash_restatement_matches_library():
    _build()
    from esn_b200 import _lib
    from esn_b200.noise import device_noise_uniforms
    lib = _lib.load()
    seed = 0xC0FFEE1234567
    u = device_noise_uniforms(seed, 3, 5, 7, first_frame=4)
    for b in range(3):
        for r in range(5):
            for n in range(7):
                assert u[b, r, n] == lib.esn_noise_uniform_host(seed, b + 4, r, n)
    big = device_noise_uniforms(11, 4, 64, 512)
    assert abs(big.mean() - 0.5) < 5e-3 and abs(big.var() - 1 / 12) < 2e-3
    assert big.min() >= 0.0 and big.max() < 1.0


def test_dropin_constructor_semantics(golden):
    from pyESN import ESN, correct_dimensions, identity
    assert correct_dimensions(None, 3) is None
    assert np.array_equal(correct_dimensions(2.0, 3), [2.0, 2.0, 2.0])
    with pytest.raises(ValueError, match="arg must have length 3"):
        correct_dimensions([1, 2], 3)
    with pytest.raises(ValueError, match="Invalid argument"):
        correct_dimensions(np.zeros((2, 2)), 2)
    assert identity(5) == 5
    with pytest.raises(Exception, match="Invalid seed"):
        ESN(2, 2, n_reservoir=8, random_state="abc")
    with pytest.raises(NotImplementedError):
        ESN(2, 2, n_reservoir=8, out_activation=np.tanh, random_state=1)
    # weights are bit-identical to the reference's for the same seed
    for name, c in cases.ESN_CASES.items():
        if c["n_res"] > 128:
            continue
        e = ESN(**cases.esn_kwargs(c))
        chk = np.array([e.W.sum(), np.abs(e.W).sum(), e.W_in.sum(), e.W_feedb.sum()])
        assert np.array_equal(chk, golden[name + "/W_checksum"])
    # falsy random_state -> numpy's global stream, as the reference (incl. 0)
    np.random.seed(123)
    a = ESN(2, 2, n_reservoir=8, random_state=None)
    np.random.seed(123)
    b = ESN(2, 2, n_reservoir=8, random_state=0)
    assert np.array_equal(a.W, b.W) and a.random_state_ is np.random.mtrand._rand
    rs = np.random.RandomState(9)
    assert ESN(2, 2, n_reservoir=8, random_state=rs).random_state_ is rs
    # predict before fit: AttributeError, as the reference
    with pytest.raises(AttributeError):
        a.predict(np.zeros((4, 2)))
    with pytest.raises(AttributeError):
        a.predict(np.zeros((4, 2)), continuation=False)


def test_spectral_radius_is_remembered_per_matrix_and_bit_identical():
    """Rebuilding an ESN from the same seed must give the same weights bit for bit as `eigvals` on every construction
    (reference libs/pyESN.py:99) -- the remembered radius only skips the repeated LAPACK call."""
    import pyESN
    pyESN._RADIUS_CACHE.clear()
    kw = dict(n_inputs=4, n_outputs=4, n_reservoir=300, spectral_radius=0.9, sparsity=0.1, random_state=42)
    a = pyESN.ESN(**kw)
    b = pyESN.ESN(**kw)
    rs = np.random.RandomState(42)
    W = rs.rand(300, 300) - 0.5
    W[rs.rand(300, 300) < 0.1] = 0
    W *= 0.9 / np.max(np.abs(np.linalg.eigvals(W)))
    assert np.array_equal(a.W, W) and np.array_equal(b.W, W)
    assert len(pyESN._RADIUS_CACHE) == 1
    c = pyESN.ESN(**dict(kw, spectral_radius=0.7))              # same draw, another radius: still one LAPACK call
    assert len(pyESN._RADIUS_CACHE) == 1 and np.allclose(c.W, W * (0.7 / 0.9), rtol=1e-15, atol=0)
    pyESN.ESN(**dict(kw, random_state=43))
    assert len(pyESN._RADIUS_CACHE) == 2


def test_helpfunc_host_parts(golden):
    from HelpFunc import HelpFunc
    for Bi in (2, 4, 6):
        assert np.allclose(HelpFunc.UnitQamConstellation(Bi), golden[f"qam/{Bi}"], rtol=0, atol=1e-15)
    with pytest.raises(TypeError):
        HelpFunc.trainMIMOESN(None, 1, 0, 6, 7, 32, 2, 2, 8, None, None)
    mag = np.exp(-np.arange(8) / 0.8)
    R = HelpFunc.ComputeChannelCorrMatrix(mag / mag.sum())
    assert R.shape == (8, 8) and np.allclose(R, R.conj().T)


def test_no_cuda_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from pyESN import ESN
    from esn_b200 import EsnB200Error
    e = ESN(2, 2, n_reservoir=8, random_state=1)
    with pytest.raises(EsnB200Error):
        e.fit(np.zeros((6, 2)), np.zeros((6, 2)))


def test_shard_range():
    from esn_b200.dist import shard_range
    for n in (0, 1, 7, 8, 1000003):
        for w in (1, 2, 3, 8):
            parts = [shard_range(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


def test_sweep_configurations_are_balanced_over_ranks():
    """cfg4: whole (size, rho, sparsity) configurations per rank, largest first; every configuration is
    placed exactly once, the plan is the same on every rank, and no rank carries more than the lightest
    one plus the largest single item."""
    sys.path.insert(0, os.path.join(ROOT, "esn-ofdm-mimo_b200"))
    from esn_b200.dist import assign_by_cost
    sizes = [64, 128, 256, 512, 1024, 2048]
    cost = [n * (n + 24) * (12.0 if n > 512 else 1.0) for n in sizes for _ in range(9)]
    for w in (1, 2, 3, 8):
        plan = assign_by_cost(cost, w)
        assert plan == assign_by_cost(list(cost), w)
        assert sorted(i for p in plan for i in p) == list(range(len(cost)))
        loads = [sum(cost[i] for i in p) for p in plan]
        assert max(loads) - min(loads) <= max(cost) + 1e-9
        for p in plan:
            assert [cost[i] for i in p] == sorted((cost[i] for i in p), reverse=True)
    assert assign_by_cost([], 4) == [[], [], [], []]


def _gloo_worker(rank, world, port, tmp):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, os.path.join(ROOT, "esn-ofdm-mimo_b200"))
    sys.path.insert(0, ROOT)
    import torch
    from esn_b200 import dist as d
    r, w, _ = d.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    # shared-readout normal equations: each rank sums its own frames, allreduce, compare
    rng = np.random.RandomState(0)
    E = rng.randn(6, 20, 9)                       # 6 frames, 20 rows, 9 columns
    D = rng.randn(6, 20, 3)
    lo, hi = d.shard_range(6, r, w)
    G = torch.from_numpy(np.einsum("brk,brl->kl", E[lo:hi], E[lo:hi])[None].copy())
    R = torch.from_numpy(np.einsum("brk,bro->ko", E[lo:hi], D[lo:hi])[None].copy())
    d.allreduce_gram_(G, R)
    Gf = np.einsum("brk,brl->kl", E, E)
    Rf = np.einsum("brk,bro->ko", E, D)
    assert np.allclose(G[0].numpy(), Gf, rtol=1e-13, atol=1e-13)
    assert np.allclose(R[0].numpy(), Rf, rtol=1e-13, atol=1e-13)
    # error counters
    cnt = torch.tensor([10 * (r + 1), r], dtype=torch.int64)
    d.allreduce_sum_(cnt)
    assert cnt.tolist() == [10 * sum(range(1, w + 1)), sum(range(w))]
    assert d.max_over_ranks(float(r), "cpu") == float(w - 1)
    # chunked shared-readout reduction: async all-reduce per chunk, partials added at the end
    red = d.GramReducer()
    for k in range(3):
        red.add(torch.full((7,), float(10 * rank + k), dtype=torch.float64))
    tot = red.finish()
    assert tot.tolist() == [float(sum(10 * r + k for r in range(w) for k in range(3)))] * 7
    assert red.bytes == 3 * 7 * 8
    d.barrier()
    open(os.path.join(tmp, f"ok{rank}"), "w").write("ok")
    torch.distributed.destroy_process_group()


def test_two_rank_gloo_allreduce(tmp_path):
    import torch.multiprocessing as mp
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_gloo_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()


def test_result_writers_reproduce_reference_files(tmp_path):
    """results_ber.csv, the calibration txt and the pickled bundle byte-for-byte / key-for-key as the
    reference's own writer lines produce them (golden made by executing those lines)."""
    import os
    import pickle
    from esn_b200 import results as R
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "soft_golden.npz"))
    eb = g["writers/EbNoDB"]
    p = R.write_results_csv(str(tmp_path / "out" / "results_ber.csv"), eb, g["writers/ue"], g["writers/um"],
                            g["writers/ce"], g["writers/cm"])
    assert open(p, "rb").read() == g["writers/csv_bytes"].tobytes()
    back = R.read_results_csv(p)
    assert back["EBN0"] == [int(x) for x in eb] and back["BER_uncoded"]["ESN"] == [float(x) for x in g["writers/ue"]]
    t = R.write_llr_calibration(str(tmp_path / "LLR_calibration_params_EbNo12dB.txt"),
                                np.stack([g["writers/a_esn"], g["writers/b_esn"]], 1),
                                np.stack([g["writers/a_mmse"], g["writers/b_mmse"]], 1))
    assert open(t, "rb").read() == g["writers/txt_bytes"].tobytes()
    ref = pickle.loads(g["writers/pkl_bytes"].tobytes())
    mine = R.results_bundle(eb, g["writers/ue"], g["writers/um"], g["writers/ce"], g["writers/cm"])
    assert mine == ref and list(mine) == list(ref)
    # uncoded-only run: coded columns are zeros, like the reference's zero-initialised arrays
    assert R.results_bundle(eb, g["writers/ue"], g["writers/um"])["BER_coded"]["ESN_calLLR"] == [0.0] * len(eb)


def test_argument_structs_have_the_layout_of_the_header(tmp_path):
    """The ctypes Structures of esn_b200/_lib.py against include/esn_b200.h compiled by gcc: same size and the same
    offset for every field (field names differ only where Python reserves them: `in` -> `inp`)."""
    import ctypes as C
    import subprocess
    from esn_b200 import _lib
    pairs = (("esn_recurrence_args", _lib.RecurrenceArgs), ("esn_tc_predict_args", _lib.TcPredictArgs),
             ("esn_tcs_args", _lib.TcsArgs))
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "esn_b200.h"', 'int main(void) {']
    for cname, cls in pairs:
        lines.append(f'  printf("{cname} size %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            cf = "in" if fname == "inp" else fname
            lines.append(f'  printf("{cname} {fname} %zu\\n", offsetof({cname}, {cf}));')
    lines += ["  return 0;", "}"]
    src, exe = tmp_path / "layout.c", tmp_path / "layout"
    src.write_text("\n".join(lines))
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split("\n")
    got = {(a, b): int(c) for a, b, c in (ln.split() for ln in out if ln.strip())}
    for cname, cls in pairs:
        assert got[(cname, "size")] == C.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert got[(cname, fname)] == getattr(cls, fname).offset, (cname, fname)
