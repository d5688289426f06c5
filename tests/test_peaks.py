"""Peak picking of the backprojection (Beamformer.find_detections / utils._detect_peaks).

CPU: the NumPy restatement (oracle/peaks_numpy.py) against vectors produced by the reference's own, unchanged code
(tests/golden/make_golden_peaks.py).  GPU: the CUDA implementation against the restatement and the vectors."""
import os

import numpy as np
import pytest

from oracle import peaks_numpy as pn

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def g():
    z = np.load(os.path.join(HERE, "golden", "peaks_golden.npz"))
    return {k: z[k] for k in z.files}


def _case(g, tag):
    return g[f"{tag}_x"], g[f"{tag}_thr"], g[f"{tag}_src"], int(g[f"{tag}_mpd"][0])


def _check_suppression_invariants(x, peaks, mpd):
    """What every valid suppression order satisfies: kept peaks are candidates more than mpd apart, and every
    dropped candidate has a kept one of at least its height within mpd."""
    cand = pn.detect_peaks(x, mpd=1)
    assert np.isin(peaks, cand).all()
    assert (np.diff(peaks) > mpd).all()
    for c in np.setdiff1d(cand, peaks):
        near = peaks[np.abs(peaks - c) <= mpd]
        assert len(near) and x[near].max() >= x[c]


@pytest.mark.parametrize("tag", ["a", "c", "d"])
def test_oracle_matches_reference_vectors(g, tag):
    x, thr, src, mpd = _case(g, tag)
    assert np.array_equal(pn.detect_peaks(x, mpd), g[f"{tag}_peaks_all"])
    pk, si = pn.find_detections(x, src, thr, mpd)
    assert np.array_equal(pk, g[f"{tag}_peaks"]) and np.array_equal(si, g[f"{tag}_sources"])


def test_oracle_on_exact_ties(g):
    """Case b: heights rounded to one decimal.  The reference's suppression order among exactly equal heights is
    whatever NumPy's unstable default sort produced on the generating host; the detections above the threshold do
    not depend on it, and both selections satisfy the suppression invariants."""
    x, thr, src, mpd = _case(g, "b")
    pk, si = pn.find_detections(x, src, thr, mpd)
    assert np.array_equal(pk, g["b_peaks"]) and np.array_equal(si, g["b_sources"])
    _check_suppression_invariants(x, pn.detect_peaks(x, mpd), mpd)
    _check_suppression_invariants(x, g["b_peaks_all"], mpd)


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["a", "b", "c", "d"])
def test_cuda_matches_oracle_and_reference_vectors(g, tag):
    from seismic_bpmf_b200 import template_search as ts

    x, thr, src, mpd = _case(g, tag)
    assert np.array_equal(ts.detect_peaks(x, mpd=mpd), pn.detect_peaks(x, mpd))
    pk, si = ts.find_detections(x, src, thr, mpd)
    assert np.array_equal(pk, g[f"{tag}_peaks"]) and np.array_equal(si, g[f"{tag}_sources"])


@pytest.mark.gpu
def test_cuda_day_length_and_degenerate_inputs():
    from seismic_bpmf_b200 import template_search as ts

    rng = np.random.default_rng(5)
    n, mpd = 4320000, 500  # one day @ 50 Hz, 10 s between events
    x = np.abs(rng.standard_normal(n)).astype(np.float32)
    x = np.convolve(x, np.ones(25, np.float32) / 25, mode="same").astype(np.float32)
    thr = np.full(n, np.float32(np.median(x) + 4 * x.std()))
    got = ts.detect_peaks(x, mpd=mpd, threshold=thr)
    want = pn.detect_peaks(x, mpd)
    want = want[x[want] > thr[want]]
    assert np.array_equal(got, want) and len(got) > 10
    for tiny in (np.zeros(0, np.float32), np.ones(2, np.float32), np.array([0, 1, 0], np.float32),
                 np.ones(50, np.float32), np.arange(50, dtype=np.float32)):
        assert np.array_equal(ts.detect_peaks(tiny, mpd=3), pn.detect_peaks(tiny, 3))
