"""Detection traces for the beamformer: drop-ins for ``envelope`` / ``envelope_parallel`` / ``saturated_envelopes``
(BPMF/template_search.py:1525-1617) and the backprojection ``time_dependent_threshold``
(BPMF/template_search.py:1418-1487), SURVEY.md 8f-f3.

Everything runs on the current CUDA device through ``libbpmf_b200_features.so`` (include/bpmf_b200_features.h);
there is no CPU fallback -- the import of the library raises when it is missing, every call raises without a device.
"""
import ctypes as C
import os

import numpy as np

from . import _lib

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BPMF_B200_FEATURES_LIB") or os.path.join(HERE, "lib", "libbpmf_b200_features.so")

fp = C.POINTER(C.c_float)
z = C.c_size_t
_SIGNATURES = {
    "bpmf_features_last_error": (C.c_char_p, []),
    "bpmf_envelope": (C.c_int, [fp, z, z, fp]),
    "bpmf_standardize_envelopes": (C.c_int, [fp, z, z, z, C.c_float, C.c_float, fp, C.POINTER(C.c_int)]),
    "bpmf_saturated_envelopes": (C.c_int, [fp, z, z, z, C.c_float, C.c_float, fp, C.POINTER(C.c_int)]),
    "bpmf_saturated_envelopes_dev": (C.c_int, [C.c_void_p, z, z, z, C.c_float, C.c_float, C.c_void_p, C.c_void_p,
                                               C.c_void_p]),
    "bpmf_bp_time_dependent_threshold": (C.c_int, [fp, z, z, z, C.c_float, C.POINTER(C.c_double)]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)
_features = None


def load():
    global _features
    if _features is not None:
        return _features
    if not os.path.exists(LIB_PATH):
        raise _lib.BpmfError("libbpmf_b200_features.so not found at %s -- build it with `python -m "
                             "seismic_bpmf_b200.build` (there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _features = lib
    return lib


def _check(rc, what):
    if rc != _lib.OK:
        msg = load().bpmf_features_last_error()
        raise _lib.BpmfError("%s failed (code %d): %s" % (what, rc, msg.decode() if msg else ""))


def envelope_parallel(traces):
    """(n_stations, n_channels, n_samples) -> modulus of the analytic signal of every trace, float32
    (template_search.py:1576-1597; the reference maps ``envelope`` over a process pool)."""
    lib = load()
    _lib.require_device()
    x = np.ascontiguousarray(traces, dtype=np.float32)
    flat = x.reshape(-1, x.shape[-1]) if x.ndim > 1 else x.reshape(1, -1)
    out = np.zeros_like(flat)
    if flat.size:
        _check(lib.bpmf_envelope(flat.ctypes.data_as(fp), flat.shape[0], flat.shape[1], out.ctypes.data_as(fp)),
               "bpmf_envelope")
    return out.reshape(x.shape)


def envelope(trace):
    """(n_samples,) -> ``np.float32(np.abs(scipy.signal.hilbert(trace)))`` (template_search.py:1599-1617)."""
    return envelope_parallel(np.asarray(trace).reshape(1, 1, -1)).reshape(-1)


def saturated_envelopes(traces, anomaly_threshold=1.0e-11, max_dynamic_range=1.0e5):
    """template_search.py:1525-1574: envelopes, standardised by the median / MAD of the present samples, zero where
    data are missing, clipped at ``max_dynamic_range``.  Returns ``(waveform_features float32 (S, C, N),
    data_availability int32 (S,))``."""
    lib = load()
    _lib.require_device()
    x = np.ascontiguousarray(traces, dtype=np.float32)
    if x.ndim != 3:
        raise ValueError("expected traces (n_stations, n_components, n_samples)")
    S, Cc, N = x.shape
    out = np.zeros_like(x)
    avail = np.zeros(S, dtype=np.int32)
    if x.size:
        _check(lib.bpmf_saturated_envelopes(x.ctypes.data_as(fp), S, Cc, N, anomaly_threshold, max_dynamic_range,
                                            out.ctypes.data_as(fp), avail.ctypes.data_as(C.POINTER(C.c_int))),
               "bpmf_saturated_envelopes")
    return out, avail


def saturated_envelopes_dev(traces, anomaly_threshold=1.0e-11, max_dynamic_range=1.0e5):
    """``saturated_envelopes`` on a float32 CUDA tensor ``(S, C, N)``: returns ``(features (S, C, N) float32,
    data_availability (S,) int32)`` as CUDA tensors, nothing crosses PCIe."""
    import torch

    lib = load()
    assert traces.is_cuda and traces.dtype == torch.float32 and traces.dim() == 3 and traces.is_contiguous()
    S, Cc, N = traces.shape
    out = torch.empty_like(traces)
    avail = torch.zeros(S, dtype=torch.int32, device=traces.device)
    _check(lib.bpmf_saturated_envelopes_dev(traces.data_ptr(), S, Cc, N, float(anomaly_threshold),
                                            float(max_dynamic_range), out.data_ptr(), avail.data_ptr(),
                                            torch.cuda.current_stream().cuda_stream), "bpmf_saturated_envelopes_dev")
    return out, avail


def standardize_envelopes(envelopes, anomaly_threshold=1.0e-11, max_dynamic_range=1.0e5):
    """The second half of ``saturated_envelopes`` on envelopes computed elsewhere (template_search.py:1544-1573)."""
    lib = load()
    _lib.require_device()
    x = np.ascontiguousarray(envelopes, dtype=np.float32)
    if x.ndim != 3:
        raise ValueError("expected envelopes (n_stations, n_components, n_samples)")
    S, Cc, N = x.shape
    out = np.zeros_like(x)
    avail = np.zeros(S, dtype=np.int32)
    if x.size:
        _check(lib.bpmf_standardize_envelopes(x.ctypes.data_as(fp), S, Cc, N, anomaly_threshold, max_dynamic_range,
                                              out.ctypes.data_as(fp), avail.ctypes.data_as(C.POINTER(C.c_int))),
               "bpmf_standardize_envelopes")
    return out, avail


def time_dependent_threshold(network_response, window, overlap=0.75, CNR_threshold=10.0):
    """Backprojection detection threshold, template_search.py:1418-1487: running median + ``CNR_threshold`` running
    MADs over windows of ``window`` samples with the given overlap, linearly interpolated to every sample
    (float64, like ``scipy.interpolate.interp1d``)."""
    lib = load()
    _lib.require_device()
    x = np.ascontiguousarray(network_response, dtype=np.float32).reshape(-1)
    window = int(window)
    shift = int((1.0 - overlap) * window)
    out = np.zeros(x.size, dtype=np.float64)
    if x.size:
        _check(lib.bpmf_bp_time_dependent_threshold(x.ctypes.data_as(fp), x.size, window, shift, CNR_threshold,
                                                    out.ctypes.data_as(C.POINTER(C.c_double))),
               "bpmf_bp_time_dependent_threshold")
    return out
