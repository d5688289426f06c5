"""Drop-in for ``beampower.beampower.beamform`` as BPMF calls it.

Reference call sites: BPMF/template_search.py:549-558 (``reduce="max"``) and :560-569
(``reduce="none"``); indirect BPMF/dataset.py:2184-2196.  Semantics: SURVEY.md Appendix A.2.
``device`` and ``num_threads`` are accepted for signature compatibility; every call computes on
the current CUDA device (no CPU fallback).
"""
import ctypes as C

import numpy as np

from .. import _lib


def _as(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


def _ptr(a, ctype):
    return a.ctypes.data_as(C.POINTER(ctype))


def _canon(waveform_features, moveouts, weights_phases, weights_sources):
    f = _as(waveform_features, np.float32)
    wp = _as(weights_phases, np.float32)
    ws = _as(weights_sources, np.float32)
    tau = np.asarray(moveouts)
    if f.ndim != 3 or tau.ndim != 3 or wp.ndim != 3 or ws.ndim != 2:
        raise ValueError("expected features (S,C,N), moveouts (K,S,P), weights_phases (S,C,P), "
                         "weights_sources (K,S)")
    S, Cc, N = f.shape
    K, S2, P = tau.shape
    if S2 != S or wp.shape != (S, Cc, P) or ws.shape != (K, S):
        raise ValueError("inconsistent shapes: features %s moveouts %s weights_phases %s weights_sources %s"
                         % (f.shape, tau.shape, wp.shape, ws.shape))
    if not np.issubdtype(tau.dtype, np.integer):
        raise TypeError("moveouts must be an integer array (samples)")
    if tau.size and (tau.max() >= 2 ** 31 or tau.min() < -2 ** 31):
        raise ValueError("moveouts exceed the int32 range")
    tau = _as(tau, np.int32)  # BPMF passes int64 (BPMF/utils.py:1270); the device kernels use int32
    return f, tau, wp, ws, (S, Cc, N, K, P)


def beamform(waveform_features, moveouts, weights_phases, weights_sources, device="gpu",
             reduce="max", out_of_bounds="strict", num_threads=None, source_offset=0):
    """Shift-and-stack network response over a grid of sources.

    ``U[s,p,t] = sum_c weights_phases[s,c,p] * waveform_features[s,c,t]``;
    ``b[k,t] = sum_s weights_sources[k,s] * sum_p U[s,p,t + moveouts[k,s,p]]``.

    Returns ``(maxbeam float32[N], maxbeam_sources int32[N])`` for ``reduce="max"`` (ties -> lowest
    source index) or ``beam float32[K,N]`` for ``reduce="none"``.
    """
    if reduce not in ("max", "none"):
        raise ValueError("reduce should be 'max' or 'none' but %r was given" % (reduce,))
    if out_of_bounds not in ("strict", "flexible"):
        raise ValueError("out_of_bounds should be 'strict' or 'flexible'")
    lib = _lib.load()
    _lib.require_device()
    f, tau, wp, ws, (S, Cc, N, K, P) = _canon(waveform_features, moveouts, weights_phases, weights_sources)
    flexible = 1 if out_of_bounds == "flexible" else 0
    if reduce == "max":
        maxbeam = np.zeros(N, dtype=np.float32)
        sources = np.zeros(N, dtype=np.int32)
        rc = lib.bpmf_beamform_max(_ptr(f, C.c_float), _ptr(tau, C.c_int), _ptr(wp, C.c_float),
                                   _ptr(ws, C.c_float), K, S, Cc, P, N, flexible, int(source_offset),
                                   _ptr(maxbeam, C.c_float), _ptr(sources, C.c_int))
        _lib.check(rc, "beamform(reduce='max')")
        return maxbeam, sources
    beam = np.zeros((K, N), dtype=np.float32)
    rc = lib.bpmf_beamform_full(_ptr(f, C.c_float), _ptr(tau, C.c_int), _ptr(wp, C.c_float),
                                _ptr(ws, C.c_float), K, S, Cc, P, N, flexible, _ptr(beam, C.c_float))
    _lib.check(rc, "beamform(reduce='none')")
    return beam


def prestack_traces(waveform_features, weights_phases, device="gpu"):
    """``U[s,p,t] = sum_c weights_phases[s,c,p] * features[s,c,t]`` (tutorial 5_backprojection.ipynb:882).

    Computed through the beam kernel with one zero-moveout, one-hot source per (station, phase).
    """
    f = _as(waveform_features, np.float32)
    wp = _as(weights_phases, np.float32)
    S, Cc, N = f.shape
    P = wp.shape[-1]
    out = np.zeros((S, P, N), dtype=np.float32)
    for p in range(P):
        wp1 = np.zeros_like(wp)
        wp1[:, :, p] = wp[:, :, p]
        beam = beamform(f, np.zeros((S, S, P), dtype=np.int32), wp1, np.eye(S, dtype=np.float32),
                        reduce="none", out_of_bounds="flexible")
        out[:, p, :] = beam
    return out
