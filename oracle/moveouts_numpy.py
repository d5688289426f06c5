"""NumPy restatement of the travel-time -> moveout conversion (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Follows BPMF/utils.py:1257-1270 (``sec_to_samp``) and BPMF/template_search.py:207-219
(``TravelTimes.get_travel_times_array(units="samples", relative_to_first=True)``).  Pinned: checked against
outputs of the reference's own functions (tests/golden/moveouts_golden.npz, made by
tests/golden/make_golden_moveouts.py in the build container).
"""
import numpy as np


def sec_to_samp(t, sr, epsilon=0.2):
    """utils.py:1265-1270 -- the arithmetic runs in the dtype of ``t`` (float32 tables stay float32)."""
    t = np.asarray(t)
    sign = np.sign(t)
    t_samp_float = abs(t * sr) + epsilon
    return np.int64(sign * np.int64(t_samp_float))


def moveouts_from_travel_times(travel_times, sampling_rate, relative_to_first=True, epsilon=0.2):
    """travel_times (K, S, P) seconds -> (moveouts int64 (K, S, P), moveout_to_tt float64 (K,)).

    template_search.py:155-165 converts each (station, phase) column with ``sec_to_samp``;
    template_search.py:212-219 subtracts the per-source minimum and keeps it, in seconds, as ``moveout_to_tt``."""
    tts = sec_to_samp(np.asarray(travel_times), sampling_rate, epsilon)
    m2t = np.zeros(tts.shape[0], dtype=np.float64)
    if relative_to_first and tts.size:
        mn = np.min(tts, axis=(1, 2), keepdims=True)
        tts = tts - mn
        m2t = mn.squeeze(axis=(1, 2)) / sampling_rate
    return tts, m2t
