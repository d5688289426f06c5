"""One day through the whole hot path: detection traces -> backprojection -> templates -> matched filter ->
detections -> inter-template CC (BASELINE.json configs[4], the order BPMF's tutorials 5 -> 8 follow).

Thin glue over the drop-in entry points of this package, written against an injectable ``backend`` so that the very
same orchestration can be driven with the CPU restatements in the tests (pipeline parity).  Array-level: the
reference's object model (``dataset.Data``, ``TemplateGroup``, pandas tables) stays in BPMF.
Reference call sites: BPMF/template_search.py:508-572 (backproject), :574-700 (find_detections),
BPMF/dataset.py:3452-3475 (template windows = pick - offset), BPMF/similarity_search.py:163-185 (set_data),
:476-546 (compute_cc_time_series), :548-724 (find_detections), BPMF/dataset.py:4776-4835 (inter-template CC).
"""
from types import SimpleNamespace

import numpy as np


def _matched_filter_detect_dev(templates, moveouts, weights, data_arr, step, window, num_dev, overlap, search_win,
                               white_noise):
    """Matched filter + rms threshold + declustering with the CC slab kept on the device: only the
    detections come back.  Same numbers as the per-template host calls it replaces."""
    import torch

    from . import device as dev

    T, S, Cc, L = templates.shape
    N = data_arr.shape[-1]
    d = torch.device("cuda", torch.cuda.current_device())
    plan = dev.MatchedFilterPlan(T, S, Cc, L, N, step=step, device=d)
    up = [torch.from_numpy(np.ascontiguousarray(x)).to(d) for x in (
        templates.astype(np.float32), moveouts.astype(np.int32), weights.astype(np.float32), data_arr.astype(np.float32))]
    plan.run(*up)
    s_, i_, _, _ = dev.detect_dev(plan.cc, torch.from_numpy(np.ascontiguousarray(white_noise, dtype=np.float32)).to(d),
                                  num_dev, window // 2, int((1.0 - overlap) * window), search_win=search_win,
                                  mode="select", n_valid=plan.n_corr, capacity=1 << 22)
    return None, [i_[s_ == t].astype(np.int64) for t in range(T)]


def gpu_backend(fused_detection=False):
    """The package's own (CUDA) entry points; raises when the libraries are missing -- no CPU fallback.
    ``fused_detection``: keep the CC slab on the device between the matched filter and the detection step."""
    from . import clib, fast_matched_filter as fmf, features, intertemplate, template_search
    from .beampower import beampower as bp

    extra = {"matched_filter_detect": _matched_filter_detect_dev} if fused_detection else {}
    return SimpleNamespace(
        **extra,
        saturated_envelopes=features.saturated_envelopes,
        bp_threshold=features.time_dependent_threshold,
        moveouts_from_travel_times=template_search.moveouts_from_travel_times,
        beamform=lambda f, mv, wp, ws: bp.beamform(f, mv, wp, ws, device="gpu", reduce="max"),
        matched_filter=lambda tp, mv, w, d, step: fmf.matched_filter(tp, mv, w, d, step, arch="gpu", check_zeros=False),
        mf_threshold=lambda cc, win, num_dev, overlap, noise: clib.time_dependent_threshold(
            cc, win, num_dev, overlap=overlap, white_noise=noise),
        select_cc_indexes=clib.select_cc_indexes,
        intertemplate_cc=intertemplate.intertemplate_cc,
        normalize_traces=normalize_traces,
    )


def normalize_traces(data_arr):
    """``MatchedFilter.set_data`` (similarity_search.py:181-185): divide every channel of the C-contiguous float32
    ``(S, C, N)`` array by its standard deviation (0 -> 1) IN PLACE on the GPU; returns the norms ``(S, C)``."""
    import ctypes as C

    from . import _lib

    lib = _lib.load()
    _lib.require_device()
    assert data_arr.dtype == np.float32 and data_arr.flags.c_contiguous and data_arr.ndim == 3
    S, Cc, N = data_arr.shape
    norm = np.ones((S, Cc), dtype=np.float32)
    fp = C.POINTER(C.c_float)
    _lib.check(lib.bpmf_normalize_traces(data_arr.ctypes.data_as(fp), S * Cc, N, norm.ctypes.data_as(fp)),
               "bpmf_normalize_traces")
    return norm


def backproject(traces, travel_times, sampling_rate, weights_phases, weights_sources, backend, window, overlap=0.75,
                cnr_threshold=10.0, search_win=None):
    """Detection traces -> moveout table -> max-beam -> time-dependent threshold -> declustered picks.

    Returns ``SimpleNamespace(moveouts, moveout_to_tt, maxbeam, sources, threshold, picks)`` where ``picks`` are the
    beam indexes (time of the earliest arrival, template_search.py:212-219) above the threshold, one per
    ``search_win`` samples, and ``sources[picks]`` the grid points they came from (template_search.py:627-639)."""
    features, availability = backend.saturated_envelopes(traces)
    moveouts, moveout_to_tt = backend.moveouts_from_travel_times(travel_times, sampling_rate, relative_to_first=True)
    maxbeam, sources = backend.beamform(features, moveouts, weights_phases, weights_sources)
    threshold = backend.bp_threshold(maxbeam, window, overlap=overlap, CNR_threshold=cnr_threshold)
    if search_win is None:
        search_win = int(moveouts.max()) + 1
    picks = np.flatnonzero(backend.select_cc_indexes(maxbeam, np.float32(threshold), search_win))
    return SimpleNamespace(features=features, availability=availability, moveouts=np.asarray(moveouts),
                           moveout_to_tt=moveout_to_tt, maxbeam=maxbeam, sources=sources, threshold=threshold,
                           picks=picks)


def extract_templates(data_arr, moveouts, phase_on_comp, picks, sources, offset, length):
    """Cut one template per pick: on component c the window starts ``offset`` samples before the arrival of phase
    ``phase_on_comp[c]`` (dataset.py:3452-3475).  Returns ``(templates (T,S,C,L), template_moveouts (T,S,C) int32,
    origin_index (T,))`` with moveouts relative to the pick index, as BPMF stores them."""
    S, Cc, N = data_arr.shape
    T = len(picks)
    templates = np.zeros((T, S, Cc, length), dtype=np.float32)
    mv = np.zeros((T, S, Cc), dtype=np.int32)
    for t, (i0, k) in enumerate(zip(picks, sources)):
        for c in range(Cc):
            mv[t, :, c] = moveouts[k, :, phase_on_comp[c]] - offset
        for s in range(S):
            for c in range(Cc):
                j = int(i0) + int(mv[t, s, c])
                lo, hi = max(j, 0), min(j + length, N)
                if hi > lo:
                    templates[t, s, c, lo - j: hi - j] = data_arr[s, c, lo:hi]
    return templates, mv, np.asarray(picks, dtype=np.int64)


def matched_filter_search(templates, template_moveouts, weights, data_arr, backend, window, num_dev=8.0, overlap=0.25,
                          search_win=None, white_noise=None, step=1):
    """CC time series of every template + rms threshold + declustering (similarity_search.py:526-533, 600-666).
    Returns ``(cc (T, n_corr), [detection indexes per template])``."""
    if white_noise is None:
        white_noise = np.random.default_rng(0).standard_normal(500).astype(np.float32)
    if search_win is None:
        search_win = templates.shape[-1]
    if hasattr(backend, "matched_filter_detect"):
        return backend.matched_filter_detect(templates, template_moveouts, weights, data_arr, step, window, num_dev,
                                             overlap, search_win, white_noise)
    cc = backend.matched_filter(templates, template_moveouts, weights, data_arr, step)
    cc[np.isnan(cc)] = 0.0  # similarity_search.py:540
    detections = []
    for t in range(cc.shape[0]):
        thr = backend.mf_threshold(cc[t], window, num_dev, overlap, white_noise)
        detections.append(np.flatnonzero(backend.select_cc_indexes(cc[t], thr, search_win)))
    return cc, detections


def full_day(traces, travel_times, sampling_rate, weights_phases, weights_sources, phase_on_comp, backend=None,
             bp_window=6000, bp_cnr=10.0, template_offset=25, template_length=200, mf_window=6000, mf_num_dev=8.0,
             max_lag=10):
    """The whole chain on one day of traces ``(S, C, N)``; see the module docstring."""
    backend = gpu_backend() if backend is None else backend
    bp_res = backproject(traces, travel_times, sampling_rate, weights_phases, weights_sources, backend, bp_window,
                         cnr_threshold=bp_cnr)
    data_arr = np.array(traces, dtype=np.float32, copy=True, order="C")
    backend.normalize_traces(data_arr)
    picks, srcs = bp_res.picks, bp_res.sources[bp_res.picks]
    templates, tmv, origin = extract_templates(data_arr, bp_res.moveouts, phase_on_comp, picks, srcs,
                                               template_offset, template_length)
    S, Cc = data_arr.shape[:2]
    weights = np.full((len(picks), S, Cc), 1.0 / (S * Cc), dtype=np.float32)
    cc, detections = (None, [])
    intertp = np.zeros((0, 0), np.float32)
    if len(picks):
        cc, detections = matched_filter_search(templates, tmv, weights, data_arr, backend, mf_window, mf_num_dev)
        intertp = backend.intertemplate_cc(templates, weights, None, max_lag=max_lag)
    return SimpleNamespace(backprojection=bp_res, templates=templates, template_moveouts=tmv, origin=origin,
                           cc=cc, detections=detections, intertemplate_cc=intertp)
