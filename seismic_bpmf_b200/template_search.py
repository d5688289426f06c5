"""Moveout tables for the beamformer.

The reference rebuilds the ``(n_sources, n_stations, n_phases)`` moveout tensor with Python loops over stations x
phases on EVERY access of ``Beamformer.moveouts`` (BPMF/template_search.py:444-454 -> ``TravelTimes.
get_travel_times_array``, :170-219) -- once per ``backproject`` call and once per detection (:636) -- after
converting seconds to samples column by column with ``utils.sec_to_samp`` (BPMF/utils.py:1257-1270).  Here the
conversion is one CUDA pass (``bpmf_moveouts_from_travel_times``, include/bpmf_b200.h) and the result is cached:
as an int32 tensor on the device for ``bpmf_beamform_*_dev`` and as a host array for callers that index it.
There is no CPU fallback.
"""
import ctypes as C

import numpy as np

from . import _lib


def moveouts_from_travel_times(travel_times, sampling_rate, relative_to_first=True, epsilon=0.2):
    """Seconds -> samples, bit-identical to the reference's ``sec_to_samp`` + ``relative_to_first``.

    travel_times : (n_sources, n_stations, n_phases) float32 or float64 -- the dtype matters: the reference
        evaluates ``|t * sr| + epsilon`` in the dtype of the table, and so does the kernel.
    Returns ``(moveouts int32 (K, S, P), moveout_to_tt float64 (K,))``; ``moveout_to_tt`` is the per-source
    first arrival in seconds (zeros when ``relative_to_first`` is False), template_search.py:212-219.
    """
    lib = _lib.load()
    _lib.require_device()
    tt = np.asarray(travel_times)
    if tt.ndim != 3:
        raise ValueError("expected travel_times (n_sources, n_stations, n_phases)")
    if tt.dtype not in (np.float32, np.float64):
        tt = tt.astype(np.float64)
    tt = np.ascontiguousarray(tt)
    K, S, P = tt.shape
    mv = np.zeros((K, S, P), dtype=np.int32)
    m2t = np.zeros(K, dtype=np.float64)
    if tt.size:
        rc = lib.bpmf_moveouts_from_travel_times(
            tt.ctypes.data_as(C.c_void_p), 1 if tt.dtype == np.float64 else 0, K, S, P, float(sampling_rate),
            float(epsilon), 1 if relative_to_first else 0, mv.ctypes.data_as(C.POINTER(C.c_int)),
            m2t.ctypes.data_as(C.POINTER(C.c_double)))
        _lib.check(rc, "bpmf_moveouts_from_travel_times")
    return mv, m2t


def moveouts_from_travel_times_dev(travel_times, sampling_rate, relative_to_first=True, epsilon=0.2):
    """The same on a float32 / float64 CUDA tensor ``(K, S, P)``: returns ``(moveouts int32 (K, S, P), moveout_to_tt
    float64 (K,))`` as CUDA tensors (what ``device.BeamformPlan`` / ``sharded.beamform_max_sharded`` take); the table
    never visits the host.  Raises if a moveout does not fit int32."""
    import torch

    lib = _lib.load()
    tt = travel_times
    assert tt.is_cuda and tt.dim() == 3 and tt.is_contiguous() and tt.dtype in (torch.float32, torch.float64)
    K, S, P = tt.shape
    mv = torch.empty((K, S, P), dtype=torch.int32, device=tt.device)
    m2t = torch.zeros(K, dtype=torch.float64, device=tt.device)
    overflow = torch.zeros(1, dtype=torch.int32, device=tt.device)
    if tt.numel():
        rc = lib.bpmf_moveouts_from_travel_times_dev(
            tt.data_ptr(), 1 if tt.dtype == torch.float64 else 0, K, S, P, float(sampling_rate), float(epsilon),
            1 if relative_to_first else 0, mv.data_ptr(), m2t.data_ptr(), overflow.data_ptr(),
            torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "bpmf_moveouts_from_travel_times_dev")
        if int(overflow.item()):
            raise _lib.BpmfError("moveouts_from_travel_times_dev: a moveout exceeds the int32 range")
    return mv, m2t


class MoveoutTable:
    """Cached moveouts of a travel-time table, keyed like ``get_travel_times_array`` is called.

    ``travel_times`` is ``{phase: {station: array(n_sources)}}`` (the layout of ``TravelTimes.travel_times``,
    a pandas table of per-(station, phase) arrays, template_search.py:104-117) or a ``(K, S, P)`` array with
    ``stations`` / ``phases`` naming its axes.  ``get(stations, phases)`` returns the same integers the
    reference's property returns, computed once per key.
    """

    def __init__(self, travel_times, sampling_rate, stations=None, phases=None, relative_to_first=True):
        self.sampling_rate = float(sampling_rate)
        self.relative_to_first = bool(relative_to_first)
        if isinstance(travel_times, np.ndarray):
            if stations is None or phases is None or travel_times.shape[1:] != (len(stations), len(phases)):
                raise ValueError("an array table needs matching `stations` and `phases` names")
            self._cols = {(sta, ph): travel_times[:, s, p] for s, sta in enumerate(stations)
                          for p, ph in enumerate(phases)}
            self.stations, self.phases = list(stations), list(phases)
        else:
            self.phases = list(travel_times.keys()) if phases is None else list(phases)
            first = travel_times[self.phases[0]]
            self.stations = list(first.keys()) if stations is None else list(stations)
            self._cols = {(sta, ph): np.asarray(travel_times[ph][sta]) for sta in self.stations for ph in self.phases}
        self._cache = {}

    def _key(self, stations, phases):
        stations = self.stations if stations is None else list(stations)
        phases = self.phases if phases is None else list(phases)
        return tuple(stations), tuple(phases)

    def get(self, stations=None, phases=None):
        """(moveouts int32 (K, S, P), moveout_to_tt float64 (K,)) for the station / phase subset."""
        key = self._key(stations, phases)
        if key not in self._cache:
            sta, ph = key
            dtype = np.result_type(*[self._cols[(s, p)].dtype for s in sta for p in ph])
            tt = np.stack([np.stack([self._cols[(s, p)] for p in ph], axis=-1) for s in sta], axis=1).astype(dtype)
            self._cache[key] = moveouts_from_travel_times(tt, self.sampling_rate, self.relative_to_first)
        return self._cache[key]

    def moveouts(self, stations=None, phases=None):
        return self.get(stations, phases)[0]

    def moveout_to_tt(self, stations=None, phases=None):
        return self.get(stations, phases)[1]

    def device_moveouts(self, stations=None, phases=None, device=None):
        """The cached table as a contiguous int32 CUDA tensor (what ``device.BeamformPlan`` takes)."""
        import torch

        key = ("dev",) + self._key(stations, phases) + (str(device),)
        if key not in self._cache:
            mv = self.moveouts(stations, phases)
            dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
            self._cache[key] = torch.from_numpy(mv).to(dev)
        return self._cache[key]


def detect_peaks(x, mpd=1, threshold=None):
    """``utils._detect_peaks(x, mpd=mpd)`` with its defaults (BPMF/utils.py:2203-2352: rising edges, suppression of
    peaks closer than ``mpd`` by decreasing height), optionally followed by ``x[peaks] > threshold[peaks]``
    (template_search.py:606-610), on the GPU.  Returns the peak indexes, increasing, int64."""
    lib = _lib.load()
    _lib.require_device()
    x = np.ascontiguousarray(x, dtype=np.float32)
    n = x.size
    thr = None
    if threshold is not None:
        thr = np.ascontiguousarray(np.broadcast_to(np.asarray(threshold, dtype=np.float32), x.shape))
    cap = max(1, n // 2 + 1)
    peaks = np.empty(cap, dtype=np.int32)
    n_peaks = C.c_size_t(0)
    fp = C.POINTER(C.c_float)
    rc = lib.bpmf_detect_peaks(x.ctypes.data_as(fp), thr.ctypes.data_as(fp) if thr is not None else None, n, int(mpd),
                               peaks.ctypes.data_as(C.POINTER(C.c_int)), cap, C.byref(n_peaks))
    _lib.check(rc, "bpmf_detect_peaks")
    return peaks[: n_peaks.value].astype(np.int64)


def find_detections(maxbeam, maxbeam_sources, detection_threshold, minimum_interevent_time_samp):
    """The peak selection of ``Beamformer.find_detections`` (BPMF/template_search.py:603-628): peaks of the composite
    network response at least ``minimum_interevent_time_samp`` apart and above the time-dependent threshold (GPU),
    then "keep the largest peak for grouped detection" and ``np.unique`` on the handful of survivors (host, the
    reference's own statements).  Returns ``(peak_indexes, source_indexes)``."""
    maxbeam = np.asarray(maxbeam)
    mpd = int(minimum_interevent_time_samp)
    peak_indexes = detect_peaks(maxbeam, mpd=mpd, threshold=detection_threshold)
    for i in range(len(peak_indexes)):
        idx = np.int32(np.arange(max(0, peak_indexes[i] - mpd / 2), min(peak_indexes[i] + mpd / 2, len(maxbeam))))
        idx_to_update = np.where(peak_indexes == peak_indexes[i])[0]
        peak_indexes[idx_to_update] = np.argmax(maxbeam[idx]) + idx[0]
    peak_indexes = np.unique(peak_indexes)
    return peak_indexes, np.asarray(maxbeam_sources)[peak_indexes]
