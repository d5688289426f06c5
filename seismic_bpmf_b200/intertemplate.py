"""Batched inter-template cross-correlation.

Replaces the loop in ``TemplateGroup.compute_intertemplate_cc`` (BPMF/dataset.py:4776-4835): the reference runs one
``fast_matched_filter.matched_filter(..., network_sum=False)`` per template -- the template's own waveforms as
"data", every other template cut by ``max_lag`` samples on both sides as "templates", zero moveouts -- then takes
the max over lags per channel and the weighted sum over channels.  Here the whole T x T matrix is one call of
``bpmf_intertemplate_cc`` (include/bpmf_b200.h); there is no CPU fallback.
"""
import ctypes as C

import numpy as np

from . import _lib


def intertemplate_cc(waveforms_arr, station_weights, pair_mask=None, max_lag=10, symmetric=True):
    """Pairwise template CCs.

    Parameters
    ----------
    waveforms_arr : (n_templates, n_stations, n_components, n_samples) float32
        ``TemplateGroup.waveforms_arr`` (BPMF/dataset.py:4786).
    station_weights : (n_templates, n_stations, n_components) float32
        Row t holds the channel weights the reference builds for template t (its ``n_stations`` closest
        stations x channel availability, dataset.py:4792-4811), normalised to sum 1.  Zero-weight channels
        are skipped.
    pair_mask : (n_templates, n_templates) bool, optional
        ``pair_mask[t, t2]`` False -> pair not computed, CC 0 (the ellipsoid-distance cut, dataset.py:4812-4814).
    max_lag : int, samples (dataset.py:4694).
    symmetric : average with the transpose as dataset.py:4833 does.

    Returns
    -------
    (n_templates, n_templates) float32
    """
    lib = _lib.load()
    _lib.require_device()
    x = np.ascontiguousarray(waveforms_arr, dtype=np.float32)
    w = np.ascontiguousarray(station_weights, dtype=np.float32)
    if x.ndim != 4 or w.shape != x.shape[:3]:
        raise ValueError("expected waveforms (T,S,C,N) and station_weights (T,S,C)")
    T, S, Cc, N = x.shape
    max_lag = int(max_lag)
    if max_lag < 0 or N <= 2 * max_lag:
        raise ValueError("max_lag=%d does not leave a template inside %d samples" % (max_lag, N))
    mask_p = None
    if pair_mask is not None:
        m = np.ascontiguousarray(np.asarray(pair_mask) != 0, dtype=np.uint8)
        if m.shape != (T, T):
            raise ValueError("pair_mask must be (T, T)")
        mask_p = m.ctypes.data_as(C.c_char_p)
    out = np.zeros((T, T), dtype=np.float32)
    if T:
        fp = C.POINTER(C.c_float)
        rc = lib.bpmf_intertemplate_cc(x.ctypes.data_as(fp), w.ctypes.data_as(fp), mask_p, T, S, Cc, N, max_lag,
                                       out.ctypes.data_as(fp))
        _lib.check(rc, "bpmf_intertemplate_cc")
    if symmetric:
        out = (out + out.T) / np.float32(2.0)
    return out


def intertemplate_cc_reference_loop(waveforms_arr, station_weights, pair_mask=None, max_lag=10, symmetric=True,
                                    matched_filter=None):
    """The reference's own loop (dataset.py:4784-4833) around a ``matched_filter`` callable (default: this
    package's GPU drop-in) -- kept for parity tests and for timing the batched kernel against it."""
    if matched_filter is None:
        from .fast_matched_filter import matched_filter
    x = np.ascontiguousarray(waveforms_arr, dtype=np.float32)
    T, S, Cc, N = x.shape
    template_arr = x[..., max_lag: N - max_lag]
    moveouts_arr = np.zeros(x.shape[:-1], dtype=np.int32)
    out = np.zeros((T, T), dtype=np.float32)
    for t in range(T):
        weights = np.repeat(np.asarray(station_weights[t], dtype=np.float32)[None], T, axis=0)
        if pair_mask is not None:
            weights[~np.asarray(pair_mask[t], dtype=bool)] = 0.0
        keep = np.sum(weights != 0.0, axis=(1, 2)) > 0
        if not keep.any():
            continue
        cc = matched_filter(template_arr[keep], moveouts_arr[keep], weights[keep], x[t], 1, network_sum=False,
                            check_zeros=False)
        out[t, keep] = np.sum(weights[keep] * np.max(cc, axis=1), axis=(-1, -2))
    if symmetric:
        out = (out + out.T) / np.float32(2.0)
    return out
