"""Mirror of the reference's ``BPMF/clib.py`` wrappers, bound to libbpmf_b200.so.

Same function names, arguments and return types as BPMF/clib.py:87-309, so code written against
``BPMF.clib`` runs unchanged; the native symbols they call (``kurtosis``,
``find_similar_moveouts[2]``, ``select_cc_indexes``, ``time_dependent_threshold``) are CUDA
implementations with the reference's C signatures (include/bpmf_b200.h group (1)).
"""
import ctypes as C
import os

import numpy as np

from . import _lib

_libc = _lib.load()
cpu_loaded = True  # name kept from BPMF/clib.py:10; the library here is the GPU one

_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int)


def kurtosis(signal, W):
    """BPMF/clib.py:87-101."""
    _lib.require_device()
    n_stations, n_components, length = signal.shape[0], signal.shape[1], signal.shape[-1]
    kurto = np.zeros(n_stations * n_components * length, dtype=np.float32)
    signal = np.float32(signal.flatten())
    _libc.kurtosis(signal.ctypes.data_as(_fp), int(W), int(n_stations), int(n_components),
                   int(length), kurto.ctypes.data_as(_fp))
    _lib.check_void("kurtosis")
    return kurto.reshape(n_stations, n_components, length)


def find_similar_sources(moveouts, source_longitude, source_latitude, cell_longitude, cell_latitude,
                         threshold, num_threads=None, num_stations_for_diff=None, method="closest"):
    """BPMF/clib.py:104-221: boolean mask of sources whose moveouts duplicate an earlier source."""
    _lib.require_device()
    n_sources, n_stations = moveouts.shape[0], moveouts.shape[1]
    n_cells_longitude = len(cell_longitude) - 1
    n_cells_latitude = len(cell_latitude) - 1
    if num_stations_for_diff is None:
        num_stations_for_diff = n_stations
    if moveouts.dtype in (np.int32, np.int64):
        print("Integer typed moveouts detected. Are you sure these are in seconds?")
    if num_threads is None:
        num_threads = int(os.environ.get("OMP_NUM_THREADS", os.cpu_count()))
    assert method in {"closest", "smallest"}, "method should be either of 'closest' or 'smallest'"
    moveouts = np.float32(moveouts.flatten())
    source_longitude = np.ascontiguousarray(source_longitude, dtype=np.float32)
    source_latitude = np.ascontiguousarray(source_latitude, dtype=np.float32)
    cell_longitude = np.ascontiguousarray(cell_longitude, dtype=np.float32)
    cell_latitude = np.ascontiguousarray(cell_latitude, dtype=np.float32)
    redundant_sources = np.zeros(n_sources, dtype=np.int32)
    fn = _libc.find_similar_moveouts2 if method == "closest" else _libc.find_similar_moveouts
    fn(moveouts.ctypes.data_as(_fp), source_longitude.ctypes.data_as(_fp),
       source_latitude.ctypes.data_as(_fp), cell_longitude.ctypes.data_as(_fp),
       cell_latitude.ctypes.data_as(_fp), np.float32(threshold), int(n_sources), int(n_stations),
       int(n_cells_longitude), int(n_cells_latitude), int(num_stations_for_diff), int(num_threads),
       redundant_sources.ctypes.data_as(_ip))
    _lib.check_void("find_similar_moveouts")
    return redundant_sources.astype(bool)


def select_cc_indexes(ccs, threshold, search_win):
    """BPMF/clib.py:224-255."""
    _lib.require_device()
    n_corr = len(ccs)
    if isinstance(threshold, (float, int)):
        threshold = threshold * np.ones(n_corr, dtype=np.float32)
    threshold = np.ascontiguousarray(threshold, dtype=np.float32)
    ccs = np.ascontiguousarray(ccs, dtype=np.float32)
    selection = np.zeros(n_corr, dtype=np.int32)
    _libc.select_cc_indexes(ccs.ctypes.data_as(_fp), threshold.ctypes.data_as(_fp), int(search_win),
                            int(n_corr), selection.ctypes.data_as(_ip))
    _lib.check_void("select_cc_indexes")
    return selection.astype(bool)


def time_dependent_threshold(time_series, sliding_window_samp, num_dev, overlap=0.66,
                             threshold_type="rms", white_noise=None, num_threads=None):
    """BPMF/clib.py:257-309 (rms threshold: mean + num_dev * std in sliding windows)."""
    _lib.require_device()
    GAUSSIAN_SAMPLE_LEN = 500
    time_series = np.array(time_series, dtype=np.float32)  # private copy: the native call mutates it
    num_samples = len(time_series)
    if white_noise is None:
        white_noise = np.random.normal(size=GAUSSIAN_SAMPLE_LEN).astype("float32")
    white_noise = np.ascontiguousarray(white_noise, dtype=np.float32)
    if num_threads is None:
        num_threads = os.cpu_count()
    half_window_samp = sliding_window_samp // 2
    shift_samp = int((1.0 - overlap) * sliding_window_samp)
    threshold = np.zeros(num_samples, dtype=np.float32)
    _libc.time_dependent_threshold(time_series.ctypes.data_as(_fp), white_noise.ctypes.data_as(_fp),
                                   float(num_dev), int(num_samples), int(half_window_samp),
                                   int(shift_samp), int(num_threads), threshold.ctypes.data_as(_fp))
    _lib.check_void("time_dependent_threshold")
    return threshold
