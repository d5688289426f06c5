"""GPU parity tests of the matched filter, through the C-ABI (host-pointer entry point) and the
device-resident entry point, against the CPU oracle on the same seeded inputs.  Tolerance: the
|dCC| < 1e-4 float32 bound of BASELINE.json's north_star, written once as TOL."""
import numpy as np
import pytest

from oracle import cpu, numpy_ref
from seismic_bpmf_b200 import _lib, fast_matched_filter as fmf
from tests import _synth

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _precisions(L, step=1, normalize="short", network_sum=True):
    lib = _lib.load()
    out = ["fp32", "auto"]
    for name in ("fp16", "fp16x3"):
        if lib.bpmf_mf_resolve_precision(L, step, 1 if normalize == "full" else 0, 1 if network_sum else 0,
                                         _lib.PRECISIONS[name]) > 0:
            out.append(name)
    return out


def _tol(prec, L=800, w=None):
    """fp32 and the hi/lo split are fp32-grade; "auto" must always meet the north-star 1e-4; an
    explicit single-pass fp16 request is held to the documented error model
    2.6e-3 * sqrt(max_t sum w^2) / sqrt(L) (include/bpmf_b200.h), which is what "auto" checks."""
    if prec == "fp32":
        return 2e-6
    if prec == "fp16x3":
        return 1e-5
    if prec == "auto":
        return TOL
    w2 = 1.0 if w is None else float((np.asarray(w, np.float64) ** 2).sum(axis=(1, 2)).max())
    return max(2e-5, 1.5 * 2.6e-3 * np.sqrt(w2) / np.sqrt(L))


def test_cfg1_shape_vs_oracle():
    """BASELINE.json configs[0]: 1 template x 4 stations x 3 channels, 1 h @ 50 Hz, L = 400."""
    tp, mv, w, d, planted = _synth.mf_case(1, 4, 3, 400, 180000)
    want = cpu.matched_filter(tp, mv, w, d, 1)
    for prec in _precisions(400):
        got = fmf.matched_filter(tp, mv, w, d, 1, arch="gpu", precision=prec, check_zeros=False)
        assert got.shape == want.shape == (1, 180000 - 400 + 1)
        assert got.dtype == np.float32 and got.flags.writeable
        err = np.abs(got - want).max()
        assert err < _tol(prec, 400, w), (prec, err)
        assert np.array_equal(got == 0, want == 0), prec
        for t, i0 in planted.items():
            assert got[t].argmax() == i0 and abs(got[t, i0] - w[t].sum()) < _tol(prec, 400, w) + 1e-6


@pytest.mark.parametrize("T,S,C,L,N", [(7, 5, 3, 200, 30011), (33, 3, 2, 800, 20000), (2, 2, 1, 37, 4133),
                                       (17, 6, 3, 800, 9000)])
def test_vs_oracle_shapes(T, S, C, L, N):
    tp, mv, w, d, _ = _synth.mf_case(T, S, C, L, N, seed=T * 131 + L, sparse_weights=(T % 2 == 1))
    want = cpu.matched_filter(tp, mv, w, d, 1)
    for prec in _precisions(L):
        got = fmf.matched_filter(tp, mv, w, d, 1, precision=prec, check_zeros=False)
        err = np.abs(got - want).max()
        assert err < _tol(prec, L, w), (prec, err)
        assert np.array_equal(got == 0, want == 0), prec


def test_float64_brute_force_agreement():
    tp, mv, w, d, _ = _synth.mf_case(3, 4, 3, 128, 12000, seed=99)
    want = numpy_ref.matched_filter(tp, mv, w, d, 1)
    for prec in _precisions(128):
        got = fmf.matched_filter(tp, mv, w, d, 1, precision=prec, check_zeros=False)
        assert np.abs(got - want).max() < _tol(prec, 128, w), prec


@pytest.mark.parametrize("step", [2, 3, 8, 25])
def test_step_is_a_subsampling(step):
    tp, mv, w, d, _ = _synth.mf_case(4, 3, 3, 100, 15000, seed=5)
    base = fmf.matched_filter(tp, mv, w, d, 1, precision="fp32", check_zeros=False)
    for prec in _precisions(100, step=step):
        got = fmf.matched_filter(tp, mv, w, d, step, precision=prec, check_zeros=False)
        assert got.shape == (4, (15000 - 100) // step + 1)
        assert np.abs(got - base[:, ::step]).max() < _tol(prec, 100, w), prec
    assert np.abs(fmf.matched_filter(tp, mv, w, d, step, precision="fp32", check_zeros=False) -
                  cpu.matched_filter(tp, mv, w, d, step)).max() < 2e-6


@pytest.mark.parametrize("step", [1, 4])
def test_normalize_full(step):
    tp, mv, w, d, _ = _synth.mf_case(3, 3, 2, 150, 9000, seed=6)
    d = d + 0.7  # non-zero mean so that "full" differs from "short"
    want = cpu.matched_filter(tp, mv, w, d, step, normalize="full")
    got = fmf.matched_filter(tp, mv, w, d, step, normalize="full", check_zeros=False)
    assert np.abs(got - want).max() < 1e-5
    short = fmf.matched_filter(tp, mv, w, d, step, check_zeros=False)
    assert np.abs(short - want).max() > 1e-3


def test_network_sum_false_layout():
    tp, mv, w, d, _ = _synth.mf_case(3, 4, 2, 64, 5000, seed=7)
    mv[:] = 0  # as BPMF/dataset.py:4778 does
    got = fmf.matched_filter(tp, mv, w, d, 1, network_sum=False, check_zeros=False)
    want = cpu.matched_filter(tp, mv, w, d, 1, network_sum=False)
    assert got.shape == (3, 5000 - 64 + 1, 4, 2)
    assert np.abs(got - want).max() < 2e-6
    summed = fmf.matched_filter(tp, mv, w, d, 1, precision="fp32", check_zeros=False)
    assert np.abs((got * w[:, None]).sum(axis=(2, 3)) - summed).max() < 1e-5


def test_known_answers_and_invariances():
    tp, mv, w, d, planted = _synth.mf_case(6, 4, 3, 256, 40000, seed=8)
    for prec in _precisions(256):
        tol = _tol(prec, 256, w)
        cc = fmf.matched_filter(tp, mv, w, d, 1, precision=prec, check_zeros=False)
        assert np.isfinite(cc).all()
        for t, i0 in planted.items():
            assert abs(cc[t, i0] - w[t].sum()) < tol + 1e-6 and cc[t].argmax() == i0
        assert np.all(np.abs(cc) <= np.abs(w).sum(axis=(1, 2))[:, None] + tol)
        assert np.abs(fmf.matched_filter(tp, mv, w, 1000.0 * d, 1, precision=prec, check_zeros=False) - cc).max() < 2 * tol
        assert np.abs(fmf.matched_filter(0.01 * tp, mv, w, d, 1, precision=prec, check_zeros=False) - cc).max() < 2 * tol
        chunks = fmf.matched_filter(tp[:, :2], mv[:, :2], w[:, :2], d[:2], 1, precision=prec, check_zeros=False) + \
            fmf.matched_filter(tp[:, 2:], mv[:, 2:], w[:, 2:], d[2:], 1, precision=prec, check_zeros=False)
        live = cc != 0
        assert np.abs(chunks - cc)[live].max() < 2 * tol  # additivity (similarity_search.py:517-538)
        zeros = fmf.matched_filter(tp, mv, w, np.zeros_like(d), 1, precision=prec, check_zeros=False)
        assert np.all(zeros == 0.0)


def test_bounds_negative_moveouts_and_zero_weight_channels():
    tp, mv, w, d, _ = _synth.mf_case(3, 3, 2, 80, 6000, seed=9, plant=False)
    mv[0] = np.arange(6).reshape(3, 2) * 11
    mv[0, 0, 0], mv[0, 1, 1] = -53, 977
    mv[2] = -5
    w[1, 2, :] = 0
    tp[1, 2, :] = np.nan  # must never be touched
    mv[1, 2, :] = 10 ** 6  # nor constrain the bounds
    want = cpu.matched_filter(tp, mv, w, d, 1)
    for prec in _precisions(80):
        got = fmf.matched_filter(tp, mv, w, d, 1, precision=prec, check_zeros=False)
        assert np.isfinite(got).all()
        assert np.all(got[0, :53] == 0) and got[0, 53] != 0
        last = 6000 - 80 - 977
        assert got[0, last] != 0 and np.all(got[0, last + 1:] == 0)
        assert np.abs(got - want).max() < _tol(prec, 80, w)


def test_edge_shapes():
    rng = np.random.default_rng(0)
    # N == L: a single correlation
    tp = rng.standard_normal((2, 1, 1, 50)).astype(np.float32)
    d = rng.standard_normal((1, 1, 50)).astype(np.float32)
    mv = np.zeros((2, 1, 1), np.int32)
    w = np.ones((2, 1, 1), np.float32)
    got = fmf.matched_filter(tp, mv, w, d, 1, check_zeros=False)
    assert got.shape == (2, 1)
    assert np.abs(got - cpu.matched_filter(tp, mv, w, d, 1)).max() < 1e-5
    # all-zero weights for one template -> a zero row (BPMF skips them, similarity_search.py:510-512)
    w[1] = 0
    got = fmf.matched_filter(tp, mv, w, d, 1, check_zeros=False)
    assert np.all(got[1] == 0)
    # empty template set
    got = fmf.matched_filter(tp[:0], mv[:0], w[:0], d, 1, check_zeros=False)
    assert got.shape == (0, 1)
    # 2-D moveouts broadcast over components; non-contiguous / float64 inputs are canonicalised
    tp3 = rng.standard_normal((2, 2, 3, 40))
    d3 = rng.standard_normal((2, 3, 900))
    mv2 = rng.integers(0, 50, (2, 2))
    w3 = np.ones((2, 2, 3))
    got = fmf.matched_filter(tp3, mv2, w3, d3[:, :, ::1], 1, check_zeros=False)
    want = cpu.matched_filter(tp3, mv2, w3, d3, 1)
    assert np.abs(got - want).max() < 2e-5  # un-normalised weights (sum = 6): auto takes the split path
    # tiny amplitudes collapse to exact zero through the stability guard
    got = fmf.matched_filter(tp3 * 1e-6, mv2, w3, d3 * 1e-6, 1, check_zeros=False)
    assert np.all(got == 0)


def test_upstream_symbol_and_check_zeros_message(capsys):
    import ctypes as C

    tp, mv, w, d, _ = _synth.mf_case(2, 3, 2, 60, 5000, seed=10)
    lib = _lib.load()
    fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int)
    n_corr = 5000 - 60 + 1
    out = np.zeros((2, n_corr), np.float32)
    sst = np.sum(tp ** 2, axis=-1).astype(np.float32)
    lib.matched_filter(tp.ctypes.data_as(fp), sst.ctypes.data_as(fp), mv.ctypes.data_as(ip), d.ctypes.data_as(fp),
                       w.ctypes.data_as(fp), 1, 60, 5000, 2, 3, 2, n_corr, 0, out.ctypes.data_as(fp))
    assert np.abs(out - cpu.matched_filter(tp, mv, w, d, 1)).max() < TOL
    d[:, :, 1000:2000] = 0
    fmf.matched_filter(tp, mv, w, d, 1, check_zeros="first")
    assert "correlation computations were skipped on the 0-th template" in capsys.readouterr().out


def test_auto_policy_and_long_templates():
    """AUTO takes the single fp16 pass only when the error model allows it; templates longer than one
    resident A image (L > 809) run as several K chunks."""
    lib = _lib.load()
    tp, mv, w, d, _ = _synth.mf_case(5, 8, 3, 800, 30000, seed=31)
    want = cpu.matched_filter(tp, mv, w, d, 1)
    got = fmf.matched_filter(tp, mv, w, d, 1, check_zeros=False)
    assert lib.bpmf_mf_last_precision() == _lib.MF_TC_FP16  # 24 equal weights, L = 800
    assert np.abs(got - want).max() < TOL
    w1 = np.zeros_like(w)
    w1[:, 0, 0] = 1.0  # a single channel carries the whole sum: single-pass error would be ~1e-4
    got = fmf.matched_filter(tp, mv, w1, d, 1, check_zeros=False)
    assert lib.bpmf_mf_last_precision() == _lib.MF_TC_FP16X3
    assert np.abs(got - cpu.matched_filter(tp, mv, w1, d, 1)).max() < 1e-5
    # both kinds in one call: AUTO decides per TEMPLATE -- the well-spread templates keep the single pass, the two
    # concentrated ones get the split, all rows land in the same slab
    wm = w.copy()
    wm[1], wm[3] = w1[1], w1[3]
    got = fmf.matched_filter(tp, mv, wm, d, 1, check_zeros=False)
    assert lib.bpmf_mf_last_precision() == _lib.MF_TC_MIXED
    want_m = cpu.matched_filter(tp, mv, wm, d, 1)
    assert np.abs(got - want_m).max() < TOL
    assert np.abs(got[[1, 3]] - want_m[[1, 3]]).max() < 1e-5
    assert np.array_equal(got == 0, want_m == 0)
    for L in (810, 1500, 2049):
        tp, mv, w, d, _ = _synth.mf_case(3, 3, 2, L, 20000, seed=L)
        want = cpu.matched_filter(tp, mv, w, d, 1)
        for prec in ("fp16", "fp16x3"):
            got = fmf.matched_filter(tp, mv, w, d, 1, precision=prec, check_zeros=False)
            assert np.abs(got - want).max() < _tol(prec, L, w), (L, prec)
            assert np.array_equal(got == 0, want == 0)


def test_large_shape_properties():
    """A BASELINE-sized slice (16 stations x 3 channels, L = 800, 20 min @ 100 Hz, 64 templates):
    too big for the float64 brute force, checked through planted detections, |cc| <= sum w, chunk
    additivity and agreement of the arithmetic variants with each other + oracle on a sub-window."""
    T, S, C, L, N = 64, 16, 3, 800, 120000
    tp, mv, w, d, planted = _synth.mf_case(T, S, C, L, N, seed=12)
    precs = _precisions(L)
    res = {p: fmf.matched_filter(tp, mv, w, d, 1, precision=p, check_zeros=False) for p in precs}
    ref = res["fp32"]
    for t, i0 in planted.items():
        assert ref[t].argmax() == i0 and abs(ref[t, i0] - 1.0) < 1e-5
    assert np.abs(ref).max() <= 1.0 + 1e-5
    for p in precs[1:]:
        assert np.abs(res[p] - ref).max() < _tol(p, L, w), p
    sub = cpu.matched_filter(tp[:4], mv[:4], w[:4], d, 1)
    assert np.abs(ref[:4] - sub).max() < 2e-6


def test_randomised_shape_sweep():
    """Seeded sweep over ragged shapes, moveout ranges (negative, beyond one tile, non multiples of 8),
    sparse weights and short traces: every arithmetic against the oracle, zeros in the same places."""
    rng = np.random.default_rng(777)
    for case in range(24):
        T = int(rng.integers(1, 40))
        S = int(rng.integers(1, 5))
        C = int(rng.integers(1, 4))
        L = int(rng.choice([8, 9, 16, 31, 64, 100, 257, 800, 820, 1000]))
        N = int(L + rng.integers(0, 6000))
        data = rng.standard_normal((S, C, N)).astype(np.float32)
        if case % 3 == 0 and N > 50:
            a = int(rng.integers(0, N - 10))
            data[:, :, a: a + int(rng.integers(1, N // 3 + 2))] = 0.0
        tp = rng.standard_normal((T, S, C, L)).astype(np.float32) * float(10.0 ** rng.integers(-3, 4))
        span = max(1, (N - L) // 2)
        mv = rng.integers(-min(span, 300), min(span, 3000) + 1, size=(T, S, C)).astype(np.int32)
        w = rng.random((T, S, C)).astype(np.float32)
        w[rng.random((T, S, C)) < 0.3] = 0.0
        if case % 4 == 1:
            w[0] = 0.0
        nrm = w.sum(axis=(1, 2), keepdims=True)
        nrm[nrm == 0] = 1.0
        w = (w / nrm).astype(np.float32)
        want = cpu.matched_filter(tp, mv, w, data, 1)
        for prec in _precisions(L):
            got = fmf.matched_filter(tp, mv, w, data, 1, precision=prec, check_zeros=False)
            assert got.shape == want.shape
            err = float(np.abs(got - want).max()) if want.size else 0.0
            assert err < _tol(prec, L, w), (case, prec, (T, S, C, L, N), err)
            assert np.isfinite(got).all()
            # exact zeros (out-of-range outputs, dead windows) coincide, up to windows whose energy sits on the
            # 1e-6 stability guard (edges of a zero gap), where float round-off may decide either way
            mismatch = np.count_nonzero((got == 0) != (want == 0))
            assert mismatch <= max(2, want.size // 500), (case, prec, mismatch)


def test_large_dynamic_range_traces():
    """A transient 1e5 times the noise level (a nearby large earthquake) must not degrade the CC of the
    quiet windows: the fp16 operands are scaled per channel by a power of two of the channel maximum and
    stay in the normal range over ~9 decades.  Checked against the float64 brute force (window-local sums):
    a global float64 prefix sum -- what mf_oracle.c and upstream use -- loses ~1e-4 of the energy of
    every window after a 3e7 spike (and a block-local one still does inside the spike's block); the
    segment-local sums of the CUDA pre-pass (mf_prepass.cu) never subtract, so they do not."""
    tp, mv, w, d, planted = _synth.mf_case(3, 3, 2, 200, 24000, seed=61, zero_gap=False)
    d = d.copy()
    d[:, :, 12000:12050] *= 1.0e5
    d[1, 0, 18000] = 3.0e7
    want = numpy_ref.matched_filter(tp, mv, w, d, 1)
    for prec in _precisions(200):
        got = fmf.matched_filter(tp, mv, w, d, 1, precision=prec, check_zeros=False)
        assert np.isfinite(got).all()
        assert np.abs(got - want).max() < _tol(prec, 200, w), prec
        n_clean = 0
        for t, i0 in planted.items():  # planted windows that the transients do not overlap still peak at sum(w)
            if abs(want[t, i0] - w[t].sum()) < 1e-6:
                n_clean += 1
                assert abs(got[t, i0] - w[t].sum()) < _tol(prec, 200, w) + 1e-6
                assert int(got[t].argmax()) == i0
        assert n_clean >= 2
    # raw counts (no normalisation, amplitudes ~1e6, the transient reaches 3e13) give the same CC in every arithmetic
    for prec in _precisions(200):
        got = fmf.matched_filter(tp * np.float32(1.0e6), mv, w, d * np.float32(1.0e6), 1, precision=prec,
                                 check_zeros=False)
        assert np.abs(got - want).max() < max(_tol(prec, 200, w), 1e-5), prec
    # tiny physical units (~1e-3): sum T^2 * sum D^2 falls under the ABSOLUTE 1e-6 stability guard and the quiet
    # windows collapse to exactly 0 by design ("try to increase the gain") -- like the oracle
    tiny_t, tiny_d = tp * np.float32(1.0e-3), d * np.float32(1.0e-3)
    want_tiny = cpu.matched_filter(tiny_t, mv, w, tiny_d, 1)
    got_tiny = fmf.matched_filter(tiny_t, mv, w, tiny_d, 1, check_zeros=False)
    assert np.count_nonzero(want_tiny == 0) > want_tiny.size // 2
    assert np.abs(got_tiny - want_tiny).max() < TOL
