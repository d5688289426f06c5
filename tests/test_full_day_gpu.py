"""One miniature day through every row of SURVEY 8: detection traces (f3) -> moveout table (f2) -> beamforming (a2)
-> backprojection threshold (f3) -> picks (a9) -> templates -> matched filter (a1) -> rms threshold + declustering
(a7, a9) -> inter-template CC (f4).  The same orchestration (seismic_bpmf_b200/workflow.py) is driven once with the
package's CUDA entry points and once with the CPU restatements: the planted events must come back, and both chains
must report the same picks and detections."""
import numpy as np
import pytest

from oracle import cpu, features_numpy as fo, libc_numpy, moveouts_numpy as mo
from seismic_bpmf_b200 import workflow
from tests import _synth
from types import SimpleNamespace



def _normalize_traces_numpy(data_arr):
    """The reference's own lines, BPMF/similarity_search.py:181-185."""
    norm = np.std(data_arr, axis=-1, keepdims=True)
    norm[norm == 0.0] = 1.0
    data_arr /= norm
    return norm.squeeze()


def _select_cc_indexes(cc, threshold, search_win):
    """libc.c:441-485, the sequential scan of oracle/libc_numpy.select_cc_indexes with the inner loop vectorised
    (same visiting order, same early exit): usable on day-long series."""
    cc = np.asarray(cc, dtype=np.float32)
    thr = np.asarray(threshold, dtype=np.float32) * np.ones(cc.size, dtype=np.float32)
    sel = np.zeros(cc.size, dtype=np.int32)
    for i in range(cc.size):
        sel[i] = 1 if cc[i] > thr[i] else 0
        start = 0 if i <= search_win else i - search_win
        gt = np.flatnonzero(cc[start:i] > cc[i])
        if gt.size:
            sel[i] = 0
            sel[start: start + gt[0]] = 0
        else:
            sel[start:i] = 0
    return sel.astype(bool)


def test_vectorised_scan_is_the_sequential_scan():
    rng = np.random.default_rng(0)
    for n, win in ((300, 7), (500, 60), (64, 100)):
        cc = rng.standard_normal(n).astype(np.float32)
        cc[rng.integers(0, n, 5)] = 2.0  # ties
        thr = (0.5 + 0.1 * rng.standard_normal(n)).astype(np.float32)
        assert np.array_equal(_select_cc_indexes(cc, thr, win), libc_numpy.select_cc_indexes(cc, thr, win).astype(bool))


def _oracle_backend():
    def beamform(f, mv, wp, ws):
        return cpu.beamform(f, mv, wp, ws)

    def mf_threshold(cc, win, num_dev, overlap, noise):
        return libc_numpy.time_dependent_threshold(np.array(cc, np.float32), noise, num_dev, win // 2,
                                                   int((1.0 - overlap) * win))[0]

    def intertp(templates, weights, mask, max_lag=10):
        from seismic_bpmf_b200 import intertemplate as itp
        return itp.intertemplate_cc_reference_loop(
            templates, weights, mask, max_lag=max_lag,
            matched_filter=lambda tp, mv, w, d, st, **kw: cpu.matched_filter(tp, mv, w, d, st, network_sum=False))

    def mv_from_tt(tt, sr, relative_to_first=True):
        mv, m2t = mo.moveouts_from_travel_times(tt, sr, relative_to_first)
        return mv.astype(np.int32), m2t

    return SimpleNamespace(
        saturated_envelopes=fo.saturated_envelopes, bp_threshold=fo.time_dependent_threshold,
        moveouts_from_travel_times=mv_from_tt, beamform=beamform,
        matched_filter=lambda tp, mv, w, d, step: cpu.matched_filter(tp, mv, w, d, step),
        mf_threshold=mf_threshold, select_cc_indexes=_select_cc_indexes, intertemplate_cc=intertp,
        normalize_traces=_normalize_traces_numpy)


@pytest.mark.gpu
def test_full_day_recovers_planted_events_and_matches_the_cpu_chain():
    case = _synth.day_case()
    args = (case["traces"], case["travel_times"], case["sr"], case["weights_phases"], case["weights_sources"],
            case["phase_on_comp"])
    gpu = workflow.full_day(*args)
    ref = workflow.full_day(*args, backend=_oracle_backend())

    truth = case["truth"]  # (source, first-arrival sample, amplitude) of the nine occurrences

    def occurrence_of(index):
        """The planted occurrence a beam / CC index belongs to: envelopes peak mid-wavelet, 20-80 samples after the
        first arrival."""
        hits = [(k, i0, amp) for k, i0, amp in truth if 20 <= int(index) - i0 <= 80]
        assert len(hits) == 1, (index, hits)
        return hits[0]

    # backprojection: every pick is a planted occurrence located at its source or a grid neighbour; the three
    # strong ones are all there
    bp = gpu.backprojection
    found = [occurrence_of(p) for p in bp.picks]
    assert len(set(found)) == len(found)
    assert {o for o in truth if o[2] > 5} <= set(found)
    for (k, i0, amp), src in zip(found, bp.sources[bp.picks]):
        assert np.linalg.norm(case["grid"][src] - case["grid"][k]) < 12.0
    assert np.array_equal(bp.picks, ref.backprojection.picks)
    assert np.array_equal(bp.sources[bp.picks], ref.backprojection.sources[ref.backprojection.picks])
    assert np.array_equal(bp.moveouts, ref.backprojection.moveouts)
    assert np.abs(bp.maxbeam - ref.backprojection.maxbeam).max() < 1e-3 * np.abs(ref.backprojection.maxbeam).max()

    # matched filter: every template re-detects the three occurrences of its own source (one strong, two weak),
    # at the same lag relative to the first arrival as the pick it was cut at
    sources_of = []
    for t, (pick, dets) in enumerate(zip(gpu.origin, gpu.detections)):
        k, i0, amp = occurrence_of(pick)
        sources_of.append(k)
        want = sorted(j0 + int(pick) - i0 for kk, j0, a in truth if kk == k)
        assert sorted(int(d) for d in dets) == want, (t, dets, want)
        assert gpu.cc[t, int(pick)] > 0.99  # the template against its own window
        assert np.array_equal(dets, ref.detections[t])
    assert np.abs(gpu.cc - ref.cc).max() < 1e-4
    # inter-template CC: unit diagonal, templates of one source resemble each other, unrelated sources do not
    itc = gpu.intertemplate_cc
    same = np.equal.outer(sources_of, sources_of)
    assert np.abs(np.diag(itc) - 1.0).max() < 1e-5
    assert itc[same].min() > 0.5 and itc[~same].max() < 0.45
    assert np.abs(itc - ref.intertemplate_cc).max() < 1e-4
    # the same day with the CC slab kept on the device between the matched filter and the detection step (f1)
    fused = workflow.full_day(*args, backend=workflow.gpu_backend(fused_detection=True))
    assert fused.cc is None and len(fused.detections) == len(gpu.detections)
    for a, b in zip(fused.detections, gpu.detections):
        assert np.array_equal(a, b)


@pytest.mark.gpu
def test_normalize_traces_matches_set_data():
    """a5: in-place division by the per-channel standard deviation, 0 -> 1 (similarity_search.py:181-185)."""
    rng = np.random.default_rng(2)
    x = (rng.standard_normal((5, 3, 40001)) * rng.uniform(1e-6, 1e4, (5, 3, 1)) + rng.uniform(-3, 3, (5, 3, 1))).astype(np.float32)
    x[2, 1] = 0.0
    x[4, 0] = 7.5  # constant trace: zero variance
    want = x.copy()
    norm_w = _normalize_traces_numpy(want)
    got = x.copy()
    norm_g = workflow.normalize_traces(got)
    x64 = x.astype(np.float64)
    std64 = x64.std(axis=-1)
    std64[std64 == 0.0] = 1.0
    assert np.abs(norm_g / std64 - 1.0).max() < 1e-6           # against float64 (NumPy's float32 std is the looser one)
    assert np.abs(norm_g / norm_w - 1.0).max() < 1e-3
    assert np.abs(got - x64 / std64[..., None]).max() < 1e-5 * np.abs(x64 / std64[..., None]).max()
    assert norm_g[2, 1] == 1.0 and np.all(got[2, 1] == 0.0)
    assert norm_g[4, 0] == 1.0 and np.all(got[4, 0] == np.float32(7.5))


def test_cpu_chain_alone_recovers_the_planted_events():
    """The orchestration itself (host logic of workflow.py), driven with the CPU restatements only: no device."""
    case = _synth.day_case()
    res = workflow.full_day(case["traces"], case["travel_times"], case["sr"], case["weights_phases"],
                            case["weights_sources"], case["phase_on_comp"], backend=_oracle_backend())
    truth = case["truth"]
    picked = {(k, i0) for p in res.backprojection.picks for k, i0, amp in truth if 20 <= int(p) - i0 <= 80}
    assert {(k, i0) for k, i0, amp in truth if amp > 5} <= picked
    for pick, dets in zip(res.origin, res.detections):
        k = next(k for k, i0, amp in truth if 20 <= int(pick) - i0 <= 80)
        lag = next(int(pick) - i0 for kk, i0, amp in truth if kk == k and 20 <= int(pick) - i0 <= 80)
        assert sorted(int(d) for d in dets) == sorted(i0 + lag for kk, i0, amp in truth if kk == k)
    assert np.abs(np.diag(res.intertemplate_cc) - 1.0).max() < 1e-5
