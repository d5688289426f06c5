"""GPU parity of the libc.so replacement symbols against golden vectors produced by the reference's
own libc.c (tests/golden/make_golden.py), through the clib mirror and the raw C-ABI."""
import numpy as np
import pytest

from oracle import libc_numpy
from seismic_bpmf_b200 import clib

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_threshold_vs_reference_golden(golden, tag):
    win, ov, nd = golden[f"thr_{tag}_params"]
    thr = clib.time_dependent_threshold(golden[f"thr_{tag}_x"], int(win), float(nd), overlap=float(ov),
                                        white_noise=golden["thr_white_noise"], num_threads=1)
    ref = golden[f"thr_{tag}_out"]
    assert thr.shape == ref.shape and thr.dtype == np.float32
    assert np.abs(thr - ref).max() < 2e-6  # float32 round-off of sums accumulated in another order


def test_threshold_mutates_input_like_reference(golden):
    import ctypes as C

    from seismic_bpmf_b200 import _lib

    lib = _lib.load()
    x = golden["thr_a_x"].copy()
    win, ov, nd = golden["thr_a_params"]
    win = int(win)
    out = np.zeros_like(x)
    fp = C.POINTER(C.c_float)
    wn = golden["thr_white_noise"]
    lib.time_dependent_threshold(x.ctypes.data_as(fp), wn.ctypes.data_as(fp), float(nd), x.size, win // 2,
                                 int((1 - ov) * win), 1, out.ctypes.data_as(fp))
    assert not np.any(x == 0)  # zeros replaced in the caller's buffer (libc.c:608-612)
    _, filled = libc_numpy.time_dependent_threshold(golden["thr_a_x"], wn, float(nd), win // 2, int((1 - ov) * win))
    assert np.abs(x - filled).max() < 1e-6


def test_threshold_overlap_clamp_and_day_length():
    rng = np.random.default_rng(1)
    n = 2159801
    x = (0.04 * rng.standard_normal(n)).astype(np.float32)
    x[500000:560000] = 0
    wn = rng.standard_normal(500).astype(np.float32)
    for ov in (0.25, 0.66):  # 0.66 is clib.py's default and reads out of bounds in the reference
        thr = clib.time_dependent_threshold(x, 45000, 8.0, overlap=ov, white_noise=wn)
        want, _ = libc_numpy.time_dependent_threshold(x, wn, 8.0, 45000 // 2, int((1 - ov) * 45000))
        assert np.isfinite(thr).all() and np.abs(thr - want).max() < 5e-6


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_select_cc_indexes_vs_reference_golden(golden, tag):
    sel = clib.select_cc_indexes(golden[f"sel_{tag}_cc"], golden[f"sel_{tag}_thr"], int(golden[f"sel_{tag}_win"][0]))
    assert sel.dtype == bool and np.array_equal(sel, golden[f"sel_{tag}_out"])


def test_select_cc_indexes_random_vs_sequential_scan():
    rng = np.random.default_rng(2)
    for n, win in [(1, 5), (2, 1), (500, 1), (4000, 37), (3000, 5000)]:
        cc = np.round(rng.standard_normal(n), 1).astype(np.float32)  # many exact ties
        thr = np.float32(0.3)
        got = clib.select_cc_indexes(cc, float(thr), win)
        want = libc_numpy.select_cc_indexes(cc, np.full(n, thr, np.float32), win).astype(bool)
        assert np.array_equal(got, want), (n, win)


def test_kurtosis_vs_reference_golden(golden):
    k = clib.kurtosis(golden["kurt_signal"], int(golden["kurt_W"][0]))
    ref = golden["kurt_out"]
    assert k.shape == ref.shape
    assert np.abs(k - ref).max() <= 2e-5 * np.abs(ref).max()
    assert np.array_equal(k == 0, ref == 0)


@pytest.mark.parametrize("tag,method", [("closest_all", "closest"), ("closest_5", "closest"), ("smallest_5", "smallest")])
def test_find_similar_sources_vs_reference_golden(golden, tag, method):
    thr, nfd = golden[f"fsm_{tag}_params"]
    red = clib.find_similar_sources(golden["fsm_moveouts"], golden["fsm_lon"], golden["fsm_lat"],
                                    golden["fsm_cell_lon"], golden["fsm_cell_lat"], float(thr), num_threads=1,
                                    num_stations_for_diff=None if nfd < 0 else int(nfd), method=method)
    assert np.array_equal(red, golden[f"fsm_{tag}_out"])


def test_find_similar_sources_multi_block():
    """More sources than one sweep block (256) so the intra-block / apply split is exercised."""
    rng = np.random.default_rng(3)
    K, S = 1500, 7
    pos = rng.uniform(0, 1, (K, 2))
    sta = rng.uniform(0, 1, (S, 2))
    tt = (np.linalg.norm(pos[:, None] - sta[None], axis=-1) * 20).astype(np.float32)
    lon, lat = pos[:, 0].astype(np.float32), pos[:, 1].astype(np.float32)
    cl = np.linspace(0, 1.001, 3).astype(np.float32)
    for method, variant, nfd in [("closest", 2, 4), ("smallest", 1, 5)]:
        got = clib.find_similar_sources(tt, lon, lat, cl, cl, 0.3, num_stations_for_diff=nfd, method=method)
        want = libc_numpy.find_similar_moveouts(tt, lon, lat, cl, cl, 0.3, nfd, variant).astype(bool)
        assert np.array_equal(got, want)
        assert 0 < got.sum() < K


def test_threshold_is_reentrant_from_a_thread_pool(golden):
    """BPMF calls time_dependent_threshold from a ThreadPoolExecutor with the GIL released
    (BPMF/similarity_search.py:590-593): 8 threads x 6 calls each on the three golden series, concurrently with
    select_cc_indexes calls, must all reproduce the reference's vectors."""
    from concurrent.futures import ThreadPoolExecutor

    g = {k: golden[k] for k in golden.files}  # NpzFile reads lazily and is not thread-safe: load everything first
    wn = g["thr_white_noise"]

    def one(k):
        tag = "abc"[k % 3]
        win, ov, nd = g[f"thr_{tag}_params"]
        thr = clib.time_dependent_threshold(g[f"thr_{tag}_x"], int(win), float(nd), overlap=float(ov),
                                            white_noise=wn, num_threads=1)
        err = float(np.abs(thr - g[f"thr_{tag}_out"]).max())
        sel = clib.select_cc_indexes(g[f"sel_{tag}_cc"], g[f"sel_{tag}_thr"], int(g[f"sel_{tag}_win"][0]))
        return err, bool(np.array_equal(sel, g[f"sel_{tag}_out"]))

    with ThreadPoolExecutor(max_workers=8) as pool:
        results = list(pool.map(one, range(48)))
    assert max(r[0] for r in results) < 2e-6
    assert all(r[1] for r in results)


def test_matched_filter_host_call_is_reentrant():
    """Two threads inside fast_matched_filter.matched_filter at once (the second one finds the per-device buffer
    cache busy and takes private buffers, csrc/abi.cu): both results equal the single-threaded ones."""
    from concurrent.futures import ThreadPoolExecutor

    from oracle import cpu
    from seismic_bpmf_b200 import fast_matched_filter as fmf
    from tests import _synth

    cases = [_synth.mf_case(9, 4, 3, 200, 40000, seed=s) for s in (1, 2, 3, 4)]
    want = [cpu.matched_filter(tp, mv, w, d, 1) for tp, mv, w, d, _ in cases]

    def one(k):
        tp, mv, w, d, _ = cases[k % 4]
        return float(np.abs(fmf.matched_filter(tp, mv, w, d, 1, precision="fp32", check_zeros=False) - want[k % 4]).max())

    with ThreadPoolExecutor(max_workers=4) as pool:
        errs = list(pool.map(one, range(16)))
    assert max(errs) < 2e-6
