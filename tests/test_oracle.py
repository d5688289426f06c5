"""CPU: the oracle itself -- C restatement vs float64 brute force, the known-answer tests of
SURVEY.md 8(c), and the libc restatement against golden vectors produced by the reference's libc.c."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import cpu, libc_numpy, numpy_ref
from tests import _synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL = 1e-4  # BASELINE.json north_star tolerance on CC


@pytest.mark.parametrize("step,normalize", [(1, "short"), (3, "short"), (1, "full"), (4, "full")])
def test_mf_c_oracle_vs_float64(step, normalize):
    tp, mv, w, d, _ = _synth.mf_case(3, 4, 3, 64, 6000, seed=1)
    a = cpu.matched_filter(tp, mv, w, d, step, normalize=normalize)
    b = numpy_ref.matched_filter(tp, mv, w, d, step, normalize=normalize)
    assert a.shape == b.shape == (3, (6000 - 64) // step + 1)
    assert np.abs(a - b).max() < 1e-5
    assert np.array_equal(a == 0, b == 0)  # exact zeros in the same places


def test_mf_no_sum_layout():
    tp, mv, w, d, _ = _synth.mf_case(2, 3, 2, 32, 2000, seed=2)
    full = cpu.matched_filter(tp, mv, w, d, 1, network_sum=False)
    assert full.shape == (2, 2000 - 32 + 1, 3, 2)
    summed = cpu.matched_filter(tp, mv, w, d, 1)
    assert np.abs((full * w[:, None]).sum(axis=(2, 3)) - summed).max() < 1e-5
    assert np.abs(full - numpy_ref.matched_filter(tp, mv, w, d, 1, network_sum=False)).max() < 1e-5


def test_mf_known_answers():
    tp, mv, w, d, planted = _synth.mf_case(4, 4, 3, 100, 20000, seed=3)
    cc = cpu.matched_filter(tp, mv, w, d, 1)
    for t, i0 in planted.items():  # KAT 1: planted template -> sum of weights
        assert abs(cc[t, i0] - w[t].sum()) < 1e-5
        assert cc[t].argmax() == i0
    assert np.all(np.abs(cc) <= np.abs(w).sum(axis=(1, 2))[:, None] + 1e-5)  # KAT 2
    assert np.abs(cpu.matched_filter(tp, mv, w, 7.0 * d, 1) - cc).max() < 1e-5  # KAT 3 scaling
    assert np.abs(cpu.matched_filter(3.0 * tp, mv, w, d, 1) - cc).max() < 1e-5
    part = cpu.matched_filter(tp[:, :2], mv[:, :2], w[:, :2], d[:2], 1) + \
        cpu.matched_filter(tp[:, 2:], mv[:, 2:], w[:, 2:], d[2:], 1)
    # KAT 4: additivity over station chunks holds wherever both chunks are in bounds
    both = (part != 0) | (cc == 0)
    assert np.abs(part - cc)[both & (cc != 0)].max() < 1e-5
    assert np.all(cpu.matched_filter(tp, mv, w, np.zeros_like(d), 1) == 0.0)  # KAT 5
    assert np.array_equal(cpu.matched_filter(tp, mv, w, d, 5), cc[:, ::5])  # KAT 6


def test_mf_bounds_and_zero_weights():
    tp, mv, w, d, _ = _synth.mf_case(2, 3, 2, 50, 3000, seed=4, plant=False)
    mv[0, 0, 0] = -37
    mv[0, 1, 1] = 411
    cc = cpu.matched_filter(tp, mv, w, d, 1)
    assert np.all(cc[0, :37] == 0) and cc[0, 37] != 0  # KAT 7
    last = 3000 - 50 - mv[0].max()
    assert cc[0, last] != 0 and np.all(cc[0, last + 1:] == 0)
    w2 = w.copy()
    w2[1, 2, :] = 0
    tp2 = tp.copy()
    tp2[1, 2, :] = np.nan  # KAT 8: zero-weight channel content is never touched
    assert np.isfinite(cpu.matched_filter(tp2, mv, w2, d, 1)).all()


def test_beam_c_oracle_vs_float64():
    f, tau, wp, ws = _synth.beam_case(40, 5, 2, 2, 3000, seed=5, sparse_sources=True)
    for oob in ("strict", "flexible"):
        mb, am = cpu.beamform(f, tau, wp, ws, out_of_bounds=oob)
        mb2, am2 = numpy_ref.beamform(f, tau, wp, ws, out_of_bounds=oob)
        assert np.abs(mb - mb2).max() < 1e-5
        full = cpu.beamform(f, tau, wp, ws, out_of_bounds=oob, reduce="none")
        assert np.abs(full - numpy_ref.beamform(f, tau, wp, ws, out_of_bounds=oob, reduce="none")).max() < 1e-5
        assert np.array_equal(full.max(axis=0), mb)  # KAT 2
        assert np.array_equal(full.argmax(axis=0), am)


def test_beam_known_answers():
    K, S, C, P, N = 12, 4, 2, 2, 2000
    f, tau, wp, ws = _synth.beam_case(K, S, C, P, N, seed=6)
    f[:] = 0
    k0, t0 = 7, 500
    for s in range(S):
        for p in range(P):
            f[s, p % C, t0 + tau[k0, s, p]] = 1.0
    mb, am = cpu.beamform(f, tau, wp, ws)
    assert abs(mb[t0] - P * ws[k0].sum()) < 1e-6 and am[t0] == k0  # KAT 1
    tau_i32 = tau.astype(np.int32)
    mb32, am32 = cpu.beamform(f, tau_i32, wp, ws)
    assert np.array_equal(mb, mb32) and np.array_equal(am, am32)  # KAT 6
    # KAT 5: ties -> lowest index (duplicate a source)
    tau2 = np.concatenate([tau, tau[k0: k0 + 1]], axis=0)
    ws2 = np.concatenate([ws, ws[k0: k0 + 1]], axis=0)
    assert cpu.beamform(f, tau2, wp, ws2)[1][t0] == k0
    # KAT 4: strict voids the tail, flexible keeps partial stacks
    g = np.ones_like(f)
    strict = cpu.beamform(g, tau, wp, ws, reduce="none")
    flex = cpu.beamform(g, tau, wp, ws, reduce="none", out_of_bounds="flexible")
    for k in range(K):
        first_bad = N - tau[k].max()
        assert np.all(strict[k, first_bad:] == 0) and strict[k, first_bad - 1] > 0
        assert flex[k, first_bad] > 0


# ---------------------------------------------------------------- libc.c golden vectors
def _thr_args(golden, tag):
    win, ov, nd = golden[f"thr_{tag}_params"]
    win = int(win)
    return golden[f"thr_{tag}_x"], golden["thr_white_noise"], float(nd), win // 2, int((1.0 - ov) * win)


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_libc_numpy_threshold_vs_reference_golden(golden, tag):
    x, wn, nd, half, shift = _thr_args(golden, tag)
    thr, filled = libc_numpy.time_dependent_threshold(x, wn, nd, half, shift)
    assert np.abs(thr - golden[f"thr_{tag}_out"]).max() < 2e-6
    assert not np.any(filled == 0)


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_libc_numpy_select_vs_reference_golden(golden, tag):
    sel = libc_numpy.select_cc_indexes(golden[f"sel_{tag}_cc"], golden[f"sel_{tag}_thr"],
                                       int(golden[f"sel_{tag}_win"][0]))
    assert np.array_equal(sel.astype(bool), golden[f"sel_{tag}_out"])


def test_libc_numpy_kurtosis_vs_reference_golden(golden):
    k = libc_numpy.kurtosis(golden["kurt_signal"], int(golden["kurt_W"][0]))
    ref = golden["kurt_out"]
    assert np.abs(k - ref).max() <= 2e-5 * np.abs(ref).max()
    assert np.array_equal(k == 0, ref == 0)


@pytest.mark.parametrize("tag,variant", [("closest_all", 2), ("closest_5", 2), ("smallest_5", 1)])
def test_libc_numpy_similar_moveouts_vs_reference_golden(golden, tag, variant):
    S = golden["fsm_moveouts"].shape[1]
    thr, nfd = golden[f"fsm_{tag}_params"]
    nfd = S if nfd < 0 else int(nfd)
    red = libc_numpy.find_similar_moveouts(golden["fsm_moveouts"], golden["fsm_lon"], golden["fsm_lat"],
                                           golden["fsm_cell_lon"], golden["fsm_cell_lat"], thr, nfd, variant)
    assert np.array_equal(red.astype(bool), golden[f"fsm_{tag}_out"])


def test_reference_libc_so_reproduces_golden(golden):
    """oracle/_ref/lib/libc.so (the reference's libc.c compiled by oracle/build.py) still gives the
    committed golden vectors -- guards against a stale or mis-built artefact."""
    path = os.path.join(ROOT, "oracle", "_ref", "lib", "libc.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref not built (no /root/reference on this box)")
    lib = C.CDLL(path)
    fp, z = C.POINTER(C.c_float), C.c_size_t
    lib.time_dependent_threshold.argtypes = [fp, fp, C.c_float, z, z, z, C.c_int, fp]
    x, wn, nd, half, shift = _thr_args(golden, "a")
    x = x.copy()
    out = np.zeros_like(x)
    lib.time_dependent_threshold(x.ctypes.data_as(fp), wn.ctypes.data_as(fp), nd, x.size, half, shift, 1,
                                 out.ctypes.data_as(fp))
    assert np.array_equal(out, golden["thr_a_out"])
