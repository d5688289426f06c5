"""ctypes front-end of oracle/liboracle.so (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Python signatures mirror the two upstream entry points BPMF calls
(BPMF/similarity_search.py:526-533, BPMF/template_search.py:549-569).
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

_lib = None
_path = None
build_kind = "portable (-march=x86-64-v3, built in the CPU container)"


def use_native():
    """Switch to a build of the same sources made on THIS host with -march=native (bench.py's CPU legs).
    Returns True when the native build is in use."""
    global _lib, _path, build_kind
    path = _build.build_oracle_native()
    if path is None:
        return False
    if path != _path:
        _lib, _path = None, path
        build_kind = "native (-O3 -fopenmp -ftree-vectorize -march=native, built on this host: %s)" % _build.cpu_model()
    return True


def lib():
    global _lib, _path
    if _lib is None:
        path = _path or os.path.join(os.path.dirname(os.path.abspath(__file__)), "liboracle.so")
        if not os.path.exists(path):
            _build.build_oracle()
        _path = path
        _lib = C.cdll.LoadLibrary(path)
        fp, ip, llp = C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_longlong)
        z = C.c_size_t
        _lib.oracle_matched_filter.argtypes = [fp, ip, fp, fp, z, z, z, z, z, z, z, C.c_int,
                                               C.c_int, C.c_int, fp]
        _lib.oracle_matched_filter.restype = C.c_int
        _lib.oracle_matched_filter_windows.argtypes = [fp, ip, fp, fp, z, z, z, z, z, z, llp, z, C.c_int, fp]
        _lib.oracle_matched_filter_windows.restype = C.c_int
        _lib.oracle_beamform.argtypes = [fp, llp, fp, fp, z, z, z, z, z, C.c_int, C.c_int, C.c_int,
                                         fp, ip, fp]
        _lib.oracle_beamform.restype = C.c_int
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def matched_filter(templates, moveouts, weights, data, step, normalize="short", network_sum=True,
                   num_threads=0):
    templates, weights, data = _f32(templates), _f32(weights), _f32(data)
    T, S, Cc, L = templates.shape
    N = data.shape[-1]
    moveouts = np.asarray(moveouts)
    if moveouts.ndim == 2:
        moveouts = np.repeat(moveouts[:, :, None], Cc, axis=2)
    moveouts = np.ascontiguousarray(moveouts, dtype=np.int32)
    n_corr = (N - L) // step + 1 if N >= L else 0
    shape = (T, n_corr) if network_sum else (T, n_corr, S, Cc)
    out = np.zeros(shape, dtype=np.float32)
    rc = lib().oracle_matched_filter(_p(templates, C.c_float), _p(moveouts, C.c_int),
                                     _p(weights, C.c_float), _p(data, C.c_float), step, L, N, T, S,
                                     Cc, n_corr, 1 if normalize == "full" else 0,
                                     1 if network_sum else 0, num_threads, _p(out, C.c_float))
    if rc != 0:
        raise MemoryError("oracle_matched_filter failed")
    return out


def matched_filter_windows(templates, moveouts, weights, data, win_start, win_len, num_threads=0):
    """cc_sum[t, win_start[w] + x] for x < win_len (step 1, "short" normalisation, network sum), without
    computing the rest of the (T, n_corr) slab -- parity checks at full day length."""
    templates, weights, data = _f32(templates), _f32(weights), _f32(data)
    T, S, Cc, L = templates.shape
    N = data.shape[-1]
    moveouts = np.ascontiguousarray(moveouts, dtype=np.int32)
    win_start = np.ascontiguousarray(win_start, dtype=np.int64)
    out = np.zeros((T, len(win_start), int(win_len)), dtype=np.float32)
    rc = lib().oracle_matched_filter_windows(_p(templates, C.c_float), _p(moveouts, C.c_int), _p(weights, C.c_float),
                                             _p(data, C.c_float), L, N, T, S, Cc, len(win_start),
                                             _p(win_start, C.c_longlong), int(win_len), num_threads,
                                             _p(out, C.c_float))
    if rc != 0:
        raise MemoryError("oracle_matched_filter_windows failed")
    return out


def beamform(waveform_features, moveouts, weights_phases, weights_sources, out_of_bounds="strict",
             reduce="max", num_threads=0):
    f, wp, ws = _f32(waveform_features), _f32(weights_phases), _f32(weights_sources)
    tau = np.ascontiguousarray(moveouts, dtype=np.int64)
    S, Cc, N = f.shape
    K, _, P = tau.shape
    flexible = 1 if out_of_bounds == "flexible" else 0
    if reduce == "max":
        mb = np.zeros(N, dtype=np.float32)
        am = np.zeros(N, dtype=np.int32)
        rc = lib().oracle_beamform(_p(f, C.c_float), _p(tau, C.c_longlong), _p(wp, C.c_float),
                                   _p(ws, C.c_float), S, Cc, P, N, K, flexible, 1, num_threads,
                                   _p(mb, C.c_float), _p(am, C.c_int), None)
        if rc != 0:
            raise MemoryError("oracle_beamform failed")
        return mb, am
    beam = np.zeros((K, N), dtype=np.float32)
    rc = lib().oracle_beamform(_p(f, C.c_float), _p(tau, C.c_longlong), _p(wp, C.c_float),
                               _p(ws, C.c_float), S, Cc, P, N, K, flexible, 0, num_threads, None,
                               None, _p(beam, C.c_float))
    if rc != 0:
        raise MemoryError("oracle_beamform failed")
    return beam
