"""NumPy restatement of the peak selection of ``Beamformer.find_detections`` (TEST INFRASTRUCTURE).

Follows BPMF/utils.py:2292-2345 (``_detect_peaks`` with its defaults: rising edges, ``mpd`` suppression by decreasing
height) and BPMF/template_search.py:603-628 (threshold, "keep the largest peak for grouped detection", unique).
Pinned: tests/golden/peaks_golden.npz was produced by the reference's own code (tests/golden/make_golden_peaks.py).

Tie rule of the suppression: the reference orders equal heights with ``np.argsort(x[ind])[::-1]`` -- NumPy's default
sort is not stable (introsort / the AVX-512 quicksort, depending on the build and the CPU), so the order of exactly
equal heights is an implementation detail of the platform the reference runs on.  This restatement, and the CUDA
kernel, use the deterministic rule of a STABLE sort in the same expression -- among equal heights the LATER peak
wins.  The committed vectors (NumPy 2.3 on an AVX-512 host) include a case with many exact ties (b: heights rounded
to one decimal) whose suppression order follows neither "earlier" nor "later"; tests/test_peaks.py checks that case
on what does not depend on the tie order (the detections above the threshold and the suppression invariants).
"""
import numpy as np


def detect_peaks(x, mpd=1):
    x = np.atleast_1d(x).astype("float64")
    n = x.size
    if n < 3:
        return np.array([], dtype=np.int64)
    nan = np.isnan(x)
    cand = np.zeros(n, dtype=bool)
    with np.errstate(invalid="ignore"):
        cand[1:-1] = (x[1:-1] > x[:-2]) & (x[2:] <= x[1:-1])
    bad = nan.copy()
    bad[1:] |= nan[:-1]
    bad[:-1] |= nan[1:]
    cand &= ~bad
    ind = np.flatnonzero(cand)
    if ind.size and mpd > 1:
        order = ind[np.argsort(x[ind], kind="stable")][::-1]  # decreasing height, ties: later peak first
        alive = np.ones(n, dtype=bool)
        keep = []
        for i in order:
            if alive[i]:
                keep.append(i)
                alive[max(0, i - mpd): i + mpd + 1] = False
        ind = np.sort(np.array(keep, dtype=np.int64))
    return ind.astype(np.int64)


def find_detections(maxbeam, maxbeam_sources, detection_threshold, mpd):
    """(peak_indexes, source_indexes) as template_search.py:603-628 computes them; ``mpd`` in samples."""
    x = np.asarray(maxbeam)
    peaks = detect_peaks(x, mpd=mpd)
    thr = np.broadcast_to(np.asarray(detection_threshold, dtype=np.float32), x.shape)
    with np.errstate(invalid="ignore"):
        peaks = peaks[x[peaks] > thr[peaks]]
    for i in range(len(peaks)):
        idx = np.int32(np.arange(max(0, peaks[i] - mpd / 2), min(peaks[i] + mpd / 2, len(x))))
        peaks[np.where(peaks == peaks[i])[0]] = np.argmax(x[idx]) + idx[0]
    peaks = np.unique(peaks)
    return peaks, np.asarray(maxbeam_sources)[peaks]
