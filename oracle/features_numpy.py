"""NumPy / SciPy restatement of the detection-trace functions (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Follows BPMF/template_search.py: ``envelope`` :1599-1617, ``saturated_envelopes`` :1525-1574 (with the process
pool of ``envelope_parallel`` :1576-1597 replaced by a loop), ``time_dependent_threshold`` :1418-1487.  Pinned:
checked against outputs of the reference's own functions (tests/golden/features_golden.npz, made by
tests/golden/make_golden_features.py in the build container).
"""
import numpy as np
from scipy.interpolate import interp1d
from scipy.signal import hilbert
from scipy.stats import median_abs_deviation as scimad


def envelope(trace):
    return np.float32(np.abs(hilbert(trace)))


def envelope_parallel(traces):
    flat = traces.reshape(-1, traces.shape[-1])
    return np.float32([envelope(t) for t in flat]).reshape(traces.shape)


def standardize_envelopes(envelopes, anomaly_threshold=1.0e-11, max_dynamic_range=1.0e5):
    """template_search.py:1544-1573, applied to given envelopes."""
    waveform_features = np.array(envelopes, dtype=np.float32, copy=True)
    n_stations, n_components, _ = waveform_features.shape
    data_availability = np.zeros(n_stations, dtype=np.int32)
    for s in range(n_stations):
        for c in range(n_components):
            missing_samples = waveform_features[s, c, :] == 0.0
            if np.sum(missing_samples) > waveform_features.shape[-1] / 2:
                waveform_features[s, c, :] = 0.0
                continue
            median = np.median(waveform_features[s, c, ~missing_samples])
            mad = scimad(waveform_features[s, c, ~missing_samples])
            if mad < anomaly_threshold:
                waveform_features[s, c, :] = 0.0
                continue
            waveform_features[s, c, :] = (waveform_features[s, c, :] - median) / mad
            waveform_features[s, c, missing_samples] = 0.0
            waveform_features[s, c, :] = np.clip(waveform_features[s, c, :], waveform_features[s, c, :],
                                                 max_dynamic_range)
            data_availability[s] += 1
    return waveform_features, data_availability


def saturated_envelopes(traces, anomaly_threshold=1.0e-11, max_dynamic_range=1.0e5):
    return standardize_envelopes(envelope_parallel(np.asarray(traces)), anomaly_threshold, max_dynamic_range)


def time_dependent_threshold(network_response, window, overlap=0.75, CNR_threshold=10.0):
    """template_search.py:1447-1487."""
    shift = int((1.0 - overlap) * window)
    n_windows = int((len(network_response) - window) // shift) + 1
    mad_ = np.zeros(n_windows + 2, dtype=np.float32)
    med_ = np.zeros(n_windows + 2, dtype=np.float32)
    time = np.zeros(n_windows + 2, dtype=np.float32)
    for i in range(1, n_windows + 1):
        i1 = i * shift
        i2 = min(network_response.size, i1 + window)
        maxbeam_window = network_response[i1:i2]
        med_[i] = np.median(maxbeam_window)
        mad_[i] = scimad(maxbeam_window)
        time[i] = (i1 + i2) / 2.0
    time[0] = 0.0
    mad_[0] = mad_[1]
    med_[0] = med_[1]
    time[-1] = len(network_response)
    mad_[-1] = mad_[-2]
    med_[-1] = med_[-2]
    threshold = med_ + CNR_threshold * mad_
    interpolator = interp1d(time, threshold, kind="slinear", fill_value=(threshold[0], threshold[-1]),
                            bounds_error=False)
    return interpolator(np.arange(0, len(network_response)))
