/* bpmf_b200_features.h -- C-ABI of libbpmf_b200_features.so: the detection traces BPMF computes right before
 * beamforming.  CUDA only (sm_100a) + cuFFT for the transforms; no CPU fallback: every entry
 * point returns BPMF_E_CUDA (include/bpmf_b200.h) without a usable device.  Host pointers in, host pointers out;
 * all arrays C-contiguous float32 unless stated. */
#ifndef BPMF_B200_FEATURES_H
#define BPMF_B200_FEATURES_H
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* message of the calling thread's last failed call */
const char *bpmf_features_last_error(void);

/* BPMF/template_search.py:1576-1617 (envelope_parallel / envelope): modulus of the analytic signal of every
 * row, |scipy.signal.hilbert(trace)| -- FFT over the whole record, like the reference. */
int bpmf_envelope(const float *traces, size_t n_traces, size_t n_samples, float *envelopes);

/* BPMF/template_search.py:1536-1574, the part of saturated_envelopes after the envelopes: per (station,
 * component) trace, samples == 0 are missing; > n_samples/2 missing or MAD < anomaly_threshold -> trace zeroed,
 * else (x - median) / MAD over the present samples (np.median / scipy median_abs_deviation in float32), missing
 * samples 0, clipped from above at max_dynamic_range.  data_availability[s] = number of valid components. */
int bpmf_standardize_envelopes(const float *envelopes, size_t n_stations, size_t n_components, size_t n_samples,
                               float anomaly_threshold, float max_dynamic_range, float *features,
                               int *data_availability);

/* BPMF/template_search.py:1525-1574 saturated_envelopes = bpmf_envelope + bpmf_standardize_envelopes without
 * leaving the device in between. */
int bpmf_saturated_envelopes(const float *traces, size_t n_stations, size_t n_components, size_t n_samples,
                             float anomaly_threshold, float max_dynamic_range, float *features,
                             int *data_availability);
/* The same with every pointer in device memory (d_features may alias d_traces), enqueued on `stream` (a cudaStream_t
 * passed as void*; synchronised once before returning): a day of traces that is already on the GPU -- the
 * backprojection that consumes the features is -- does not make a round trip through host memory. */
int bpmf_saturated_envelopes_dev(const float *d_traces, size_t n_stations, size_t n_components, size_t n_samples,
                                 float anomaly_threshold, float max_dynamic_range, float *d_features,
                                 int *d_data_availability, void *stream);

/* BPMF/template_search.py:1418-1487 time_dependent_threshold (backprojection): windows i = 1..n_windows of
 * `window` samples every `shift` = int((1 - overlap) * window) samples, median + cnr_threshold * MAD per window
 * (float32), linear interpolation through (window centre, value) with the first / last value held to the ends,
 * evaluated at every sample in float64. */
int bpmf_bp_time_dependent_threshold(const float *network_response, size_t n_samples, size_t window, size_t shift,
                                     float cnr_threshold, double *threshold);

#ifdef __cplusplus
}
#endif
#endif
