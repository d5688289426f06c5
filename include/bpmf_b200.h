/*
 * bpmf_b200.h -- C-ABI of libbpmf_b200.so, the B200 (sm_100a) drop-in for the native hot path of
 * ebeauce/Seismic_BPMF.  Plain pointers and sizes only; no torch / C++ types cross this boundary.
 *
 * Three groups of entry points, each citing the reference interface it replaces
 * (paths relative to the reference tree):
 *
 *  (1) The symbols of BPMF/lib/libc.so that BPMF/clib.py binds (clib.py:26-84, libc.h:1-11,
 *      libc.c:11,55,225,389,413,421,441,488,502,516).  Same names, same signatures, same
 *      ownership rules (caller-allocated numpy buffers, void return, input of
 *      time_dependent_threshold is mutated) -- so `clib.py` works unchanged when this library is
 *      installed as BPMF/lib/libc.so.  Work runs on the current CUDA device.
 *  (2) The native entry points behind `fast_matched_filter.matched_filter`
 *      (BPMF/similarity_search.py:526-533, BPMF/dataset.py:4818-4827) and
 *      `beampower.beampower.beamform` (BPMF/template_search.py:549-569).  Those two packages are
 *      un-vendored pip dependencies of the reference (pyproject.toml:28-29); their C symbol names
 *      are reproduced from the upstream projects as best known ([UPSTREAM-UNVERIFIED], SURVEY.md
 *      Appendix A) next to explicit `bpmf_*` entry points.
 *  (3) `caract_GPU` (BPMF/GPU.cu:7-24).
 *
 * Every `bpmf_*` function returns 0 on success or a negative BPMF_E_* code; the message is
 * available from bpmf_last_error() (thread-local).  There is NO CPU fallback: without a usable
 * CUDA device every compute entry point fails with BPMF_E_CUDA.
 */
#ifndef BPMF_B200_H
#define BPMF_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BPMF_OK 0
#define BPMF_E_CUDA (-1)     /* CUDA runtime error / no device */
#define BPMF_E_ARG (-2)      /* invalid argument / unsupported shape */
#define BPMF_E_NOMEM (-3)    /* host or device allocation failed */

/* matched-filter arithmetic selector */
#define BPMF_MF_AUTO 0       /* tensor path when the shape allows it, else fp32 */
#define BPMF_MF_FP32 1       /* direct form on FP32 CUDA cores (exact-grade, ~1e-6) */
#define BPMF_MF_TC_FP16 2    /* tcgen05, fp16 operands (power-of-two pre-scaled), fp32 accumulate */
#define BPMF_MF_TC_FP16X3 3  /* tcgen05, hi/lo split of both operands (3 MMAs), fp32-grade */
#define BPMF_MF_TC_MIXED 4   /* reported only (bpmf_mf_last_precision): AUTO ran FP16 for the templates whose
                                weights allow it and FP16X3 for the others */

const char *bpmf_last_error(void);
/* The reference-named symbols of group (1) return void (BPMF/libc.h:1-11) and can only report a
 * failure on stderr.  bpmf_last_status() returns the code (0 or BPMF_E_*) of the calling thread's
 * last such call, so that a binding can raise; on failure the float outputs (threshold, kurtosis)
 * are additionally filled with NaN -- `cc > NaN` is false, so a caller that cannot check does not
 * mistake a failed threshold for "everything is a detection". */
int bpmf_last_status(void);
int bpmf_device_count(void);
/* library build id, e.g. "bpmf_b200 r1 sm_100a" */
const char *bpmf_version(void);
/* number of kernel launches issued by this library since process start (all threads) */
unsigned long long bpmf_launch_count(void);

/* The persistent matched-filter kernel occupies every SM (one CTA each, all of its shared memory).
 * A caller that overlaps a collective (NCCL all-gather of finished CC slabs) with the next batch
 * reserves a few SMs for it; 0 (default) = use them all. */
void bpmf_set_reserved_sms(int n);

/* Optional instrumentation for bench.py: while enabled, the dominant kernel of every
 * bpmf_matched_filter_dev / bpmf_beamform_*_dev call made by the calling thread is bracketed by
 * CUDA events on the launching stream (up to 256 launches).  bpmf_main_kernel_times() waits for
 * them, writes their durations in milliseconds and returns how many; it also clears the list. */
void bpmf_set_kernel_timing(int enable);
int bpmf_main_kernel_times(float *ms, int capacity);

/* ---------------------------------------------------------------- (3) BPMF/GPU.cu:7-24 */
void caract_GPU(void);

/* ---------------------------------------------------------------- (2a) matched filter
 * cc_sum[t, n] = sum_{s,c: w!=0} w[t,s,c] * CC(templates[t,s,c,:], data[s,c, n*step+moveouts[t,s,c] : +L])
 * Layouts (C order): templates (T,S,C,L) f32; moveouts (T,S,C) i32 (may be negative);
 * weights (T,S,C) f32; data (S,C,N) f32; n_corr = (N-L)/step + 1.
 * normalize: 0 "short" (no window-mean removal), 1 "full".
 * network_sum: 1 -> out (T, n_corr); 0 -> out (T, n_corr, S, C) unweighted per-channel CC.
 */
size_t bpmf_mf_n_corr(size_t n_samples_data, size_t n_samples_template, size_t step);

/* Host-pointer entry point (what a ctypes/cffi binding calls).  Copies in, runs on the current
 * device, copies out; templates are processed in chunks sized to the free device memory. */
int bpmf_matched_filter(const float *templates, const int *moveouts, const float *weights,
                        const float *data, size_t step, size_t n_samples_template,
                        size_t n_samples_data, size_t n_templates, size_t n_stations,
                        size_t n_components, int normalize, int network_sum, int precision,
                        float *cc_out);

/* Device-pointer entry point: every pointer is device memory, work is enqueued on `stream`
 * (a cudaStream_t passed as void*), no host synchronisation for an explicit precision.  `workspace` must hold at least
 * bpmf_mf_workspace_bytes(...) bytes.  cc_pitch = elements between consecutive template rows of
 * d_cc (>= n_corr; ignored when network_sum == 0).  The tensor-core arithmetics (BPMF_MF_TC_*, and
 * BPMF_MF_AUTO whenever the shape allows them) accumulate with vector reductions: d_cc must be
 * 16-byte aligned and cc_pitch a multiple of 4 elements, otherwise the call returns BPMF_E_ARG
 * (the Python wrappers pad rows to 64 elements). */
size_t bpmf_mf_workspace_bytes(size_t n_samples_template, size_t n_samples_data,
                               size_t n_templates, size_t n_stations, size_t n_components,
                               size_t step, int normalize, int network_sum, int precision);
int bpmf_matched_filter_dev(const float *d_templates, const int *d_moveouts,
                            const float *d_weights, const float *d_data, size_t step,
                            size_t n_samples_template, size_t n_samples_data, size_t n_templates,
                            size_t n_stations, size_t n_components, int normalize,
                            int network_sum, int precision, void *d_workspace,
                            size_t workspace_bytes, float *d_cc, size_t cc_pitch, void *stream);
/* The same work in two steps, for callers that run many template batches against the same traces
 * (BPMF/similarity_search.py:783-790 loops template chunks over one day of data): the trace-side
 * staging (fp16 copies, window energies) is prepared once into `d_trace_ws` and reused, read-only,
 * by every bpmf_mf_run_dev call -- also concurrently from several streams, each with its own
 * `d_template_ws`.  L, N, S, C, step, normalize, network_sum and precision must match. */
size_t bpmf_mf_trace_workspace_bytes(size_t n_samples_template, size_t n_samples_data, size_t n_stations,
                                     size_t n_components, size_t step, int normalize, int network_sum,
                                     int precision);
size_t bpmf_mf_template_workspace_bytes(size_t n_samples_template, size_t n_samples_data,
                                        size_t n_templates, size_t n_stations, size_t n_components,
                                        size_t step, int normalize, int network_sum, int precision);
int bpmf_mf_prepare_traces_dev(const float *d_data, size_t step, size_t n_samples_template,
                               size_t n_samples_data, size_t n_stations, size_t n_components,
                               int normalize, int network_sum, int precision, void *d_trace_ws,
                               size_t trace_ws_bytes, void *stream);
int bpmf_mf_run_dev(const float *d_templates, const int *d_moveouts, const float *d_weights,
                    const float *d_data, size_t step, size_t n_samples_template, size_t n_samples_data,
                    size_t n_templates, size_t n_stations, size_t n_components, int normalize,
                    int network_sum, int precision, void *d_trace_ws, size_t trace_ws_bytes,
                    void *d_template_ws, size_t template_ws_bytes, float *d_cc, size_t cc_pitch,
                    void *stream);
/* frees the device buffers the host-pointer entry points keep cached on the current device */
void bpmf_release_cache(void);
/* BPMF loops template chunks over one day of traces, one matched_filter call per chunk on the same `data` array
 * (BPMF/similarity_search.py:773-790).  Between hold and release the caller promises that the host buffer at `data`
 * does not change: bpmf_matched_filter then uploads and stages it once and reuses the device copy for every call
 * that passes the same pointer and shape (same L, step, normalize, precision).  Nothing is cached without a hold. */
int bpmf_mf_hold_traces(const float *data, size_t n_samples_data, size_t n_stations, size_t n_components);
void bpmf_mf_release_traces(void);

/* Shape-only check of a precision selector: returns the arithmetic it maps to, or a negative value
 * when it is not available for the shape.  For BPMF_MF_AUTO this is the fastest arithmetic the
 * shape allows; the call itself classifies every TEMPLATE from its weights (8 bytes read back, the
 * one host synchronisation of bpmf_matched_filter_dev): a template takes BPMF_MF_TC_FP16 when the
 * predicted rounding error of its CC sum, 2.6e-3 * sqrt(sum_ch w^2) / sqrt(L), is below 6e-5,
 * otherwise BPMF_MF_TC_FP16X3; a call with both kinds runs the two classes one after the other
 * (BPMF_MF_TC_MIXED).  bpmf_mf_last_precision() reports what the calling thread's last call used. */
int bpmf_mf_resolve_precision(size_t n_samples_template, size_t step, int normalize,
                              int network_sum, int precision);
int bpmf_mf_last_precision(void);

/* Upstream fast_matched_filter CPU/GPU symbol [UPSTREAM-UNVERIFIED signature]: the function
 * fast_matched_filter's own ctypes wrapper binds.  sum_square_templates is accepted and ignored
 * (recomputed on the device). */
void matched_filter(float *templates, float *sum_square_templates, int *moveouts, float *data,
                    float *weights, size_t step, size_t n_samples_template, size_t n_samples_data,
                    size_t n_templates, size_t n_stations, size_t n_components, size_t n_corr,
                    int normalize, float *cc_sum);
void matched_filter_no_sum(float *templates, float *sum_square_templates, int *moveouts,
                           float *data, float *weights, size_t step, size_t n_samples_template,
                           size_t n_samples_data, size_t n_templates, size_t n_stations,
                           size_t n_components, size_t n_corr, int normalize, float *cc);

/* ---------------------------------------------------------------- (2b) beamforming
 * U[s,p,t] = sum_c weights_phases[s,c,p] * features[s,c,t]
 * b[k,t]   = sum_s weights_sources[k,s] * sum_p U[s,p,t + moveouts[k,s,p]]
 * Layouts: features (S,C,N) f32; moveouts (K,S,P) i32; weights_phases (S,C,P) f32;
 * weights_sources (K,S) f32.  flexible: 0 "strict", 1 "flexible" (template_search.py:529-537).
 * source_offset is added to every reported argmax (source-sharded callers).
 */
int bpmf_beamform_max(const float *features, const int *moveouts, const float *weights_phases,
                      const float *weights_sources, size_t n_sources, size_t n_stations,
                      size_t n_components, size_t n_phases, size_t n_samples, int flexible,
                      int source_offset, float *maxbeam, int *maxbeam_sources);
int bpmf_beamform_full(const float *features, const int *moveouts, const float *weights_phases,
                       const float *weights_sources, size_t n_sources, size_t n_stations,
                       size_t n_components, size_t n_phases, size_t n_samples, int flexible,
                       float *beam);
size_t bpmf_beam_workspace_bytes(size_t n_sources, size_t n_stations, size_t n_components,
                                 size_t n_phases, size_t n_samples);
/* device-pointer variants, enqueued on `stream`, no host synchronisation */
int bpmf_beamform_max_dev(const float *d_features, const int *d_moveouts,
                          const float *d_weights_phases, const float *d_weights_sources,
                          size_t n_sources, size_t n_stations, size_t n_components,
                          size_t n_phases, size_t n_samples, int flexible, int source_offset,
                          void *d_workspace, size_t workspace_bytes, float *d_maxbeam,
                          int *d_maxbeam_sources, void *stream);
int bpmf_beamform_full_dev(const float *d_features, const int *d_moveouts,
                           const float *d_weights_phases, const float *d_weights_sources,
                           size_t n_sources, size_t n_stations, size_t n_components,
                           size_t n_phases, size_t n_samples, int flexible, void *d_workspace,
                           size_t workspace_bytes, float *d_beam, void *stream);
/* merge per-shard (max, argmax) pairs: in-place  a = better(a, b), ties -> lower source index */
int bpmf_beam_merge_dev(float *d_max_a, int *d_arg_a, const float *d_max_b, const int *d_arg_b,
                        size_t n_samples, void *stream);

/* ---------------------------------------------------------------- (2c) batched inter-template CC
 * Replaces the loop of T fast_matched_filter calls (network_sum=False, zero moveouts, one template's
 * waveforms as data) in TemplateGroup.compute_intertemplate_cc, BPMF/dataset.py:4784-4830:
 *   cc[t, t2] = sum_{s,c} station_weights[t,s,c] * max over the 2*max_lag+1 lags of
 *               CC( waveforms[t2,s,c, max_lag : n_samples-max_lag],  waveforms[t,s,c, lag : ...] )
 * waveforms (T,S,C,n_samples) f32; station_weights (T,S,C) f32 (dataset.py:4806-4811, already normalised);
 * pair_mask (T,T) bytes, 0 = pair not computed (cc = 0; dataset.py:4812-4814), NULL = all pairs; cc (T,T) f32,
 * NOT symmetrised (the caller averages with the transpose, dataset.py:4833). */
int bpmf_intertemplate_cc(const float *waveforms, const float *station_weights,
                          const unsigned char *pair_mask, size_t n_templates, size_t n_stations,
                          size_t n_components, size_t n_samples, size_t max_lag, float *cc);
size_t bpmf_intertemplate_workspace_bytes(size_t n_templates, size_t n_stations, size_t n_components,
                                          size_t max_lag);
int bpmf_intertemplate_cc_dev(const float *d_waveforms, const float *d_station_weights,
                              const unsigned char *d_pair_mask, size_t n_templates, size_t n_stations,
                              size_t n_components, size_t n_samples, size_t max_lag, void *d_workspace,
                              size_t workspace_bytes, float *d_cc, void *stream);

/* ---------------------------------------------------------------- (2d) travel times -> moveouts
 * sec_to_samp (BPMF/utils.py:1257-1270): sign(t) * int64(|t * sampling_rate| + epsilon), evaluated in the dtype
 * of the table (is_float64: 0 = float32, 1 = float64), then, if relative_to_first, minus the per-source minimum
 * over (station, phase) (BPMF/template_search.py:212-219); moveout_to_tt[k] = that minimum / sampling_rate
 * (may be NULL).  travel_times (K,S,P); moveouts (K,S,P) int32 -- the layout bpmf_beamform_* consume.  The
 * device variant raises *d_overflow when a value does not fit int32 (the host variant returns BPMF_E_ARG). */
int bpmf_moveouts_from_travel_times(const void *travel_times, int is_float64, size_t n_sources,
                                    size_t n_stations, size_t n_phases, double sampling_rate, double epsilon,
                                    int relative_to_first, int *moveouts, double *moveout_to_tt);
int bpmf_moveouts_from_travel_times_dev(const void *d_travel_times, int is_float64, size_t n_sources,
                                        size_t n_stations, size_t n_phases, double sampling_rate,
                                        double epsilon, int relative_to_first, int *d_moveouts,
                                        double *d_moveout_to_tt, int *d_overflow, void *stream);

/* ---------------------------------------------------------------- (2e) trace normalisation
 * MatchedFilter.set_data, BPMF/similarity_search.py:181-185: every one of the n_channels (station, component) traces
 * is divided IN PLACE by its population standard deviation (0 -> 1); norm[n_channels] receives the divisors
 * (self.data_norm).  Sums in float64.  The device variant needs 2 * n_channels doubles of scratch. */
int bpmf_normalize_traces(float *data, size_t n_channels, size_t n_samples, float *norm);
int bpmf_normalize_traces_dev(float *d_data, size_t n_channels, size_t n_samples, float *d_norm,
                              double *d_scratch, void *stream);

/* ---------------------------------------------------------------- (1) BPMF/lib/libc.so symbols */
/* libc.c:11-16   ; clib.py:26-33 */
void kurtosis(float *signal, int W, int n_stations, int n_components, int length, float *kurto);
/* libc.c:55-68   ; clib.py:35-49 */
void find_similar_moveouts(float *moveouts, float *source_longitude, float *source_latitude,
                           float *cell_longitude, float *cell_latitude, float threshold,
                           size_t n_sources, size_t n_stations, size_t n_cells_longitude,
                           size_t n_cells_latitude, size_t n_stations_for_diff, int num_threads,
                           int *redundant_sources);
/* libc.c:225-238 ; clib.py:51-65 */
void find_similar_moveouts2(float *moveouts, float *source_longitude, float *source_latitude,
                            float *cell_longitude, float *cell_latitude, float threshold,
                            size_t n_sources, size_t n_stations, size_t n_cells_longitude,
                            size_t n_cells_latitude, size_t n_stations_for_diff, int num_threads,
                            int *redundant_sources);
/* libc.c:441-445 ; clib.py:67-73 */
void select_cc_indexes(float *cc, float *threshold, size_t search_win, size_t n_corr,
                       int *selection);
/* libc.c:516-524 ; clib.py:75-84.  time_series is overwritten (zeros replaced), like libc.c:610 */
void time_dependent_threshold(float *time_series, float *gaussian_sample, float num_dev,
                              size_t num_samples, size_t half_window_samp, size_t shift_samp,
                              int num_threads, float *threshold);
/* helper symbols the reference library also exports (libc.h:5-10); tiny host utilities */
void swap(float *xp, float *yp);
void selectionSort(float *arr, int n);
void sort_indexes(const float *moveouts, size_t *sorted_indexes, size_t n_stations);
float _compute_average(float *x, size_t len_x);
float _compute_std(float *x, float x_mean, size_t len_x);

/* device-resident batch form of time_dependent_threshold: B series of n samples each (row pitch
 * `pitch` elements), e.g. the (T, n_corr) CC slab straight after bpmf_matched_filter_dev, so the
 * CC never leaves the GPU before thresholding.  Series are modified in place. */
size_t bpmf_threshold_workspace_bytes(size_t n_series, size_t num_samples, size_t half_window_samp,
                                      size_t shift_samp);
int bpmf_time_dependent_threshold_dev(float *d_series, size_t pitch, size_t n_series,
                                      const float *d_gaussian_sample, float num_dev,
                                      size_t num_samples, size_t half_window_samp,
                                      size_t shift_samp, void *d_workspace, size_t workspace_bytes,
                                      float *d_threshold, size_t threshold_pitch, void *stream);

/* Candidate detections without moving the CC slab off the GPU (replaces the
 * per-template download + time_dependent_threshold + peak selection of
 * BPMF/similarity_search.py:600-666).  For each of the n_series rows: rms threshold as above (the rows
 * are NOT modified), then every sample with cc > threshold (mode 0) that also survives the
 * declustering of select_cc_indexes with `search_win` (mode 1; libc.c:441-485) is appended, unordered,
 * to the compact lists (series, index, cc, threshold).  *d_count receives the number of hits, which
 * may exceed `capacity` (then only `capacity` records were written).  Workspace:
 * bpmf_threshold_workspace_bytes(). */
int bpmf_detect_dev(const float *d_series, size_t pitch, size_t n_series, const float *d_gaussian_sample,
                    float num_dev, size_t num_samples, size_t half_window_samp, size_t shift_samp,
                    size_t search_win, int mode, void *d_workspace, size_t workspace_bytes,
                    unsigned *d_count, int *d_series_idx, int *d_index, float *d_value, float *d_thr_value,
                    size_t capacity, void *stream);
/* The same with the two remaining per-template steps of MatchedFilter.find_detections folded in:
 *   d_thr_cap (n_series floats or NULL): threshold = min(cap, threshold), the caller passing
 *     max_CC_threshold * sum(weights) per template (BPMF/similarity_search.py:629);
 *   use_kurtosis / max_kurtosis: a row whose excess kurtosis (scipy.stats.kurtosis defaults, all
 *     num_samples samples, float64) exceeds max_kurtosis detects nothing -- the reference zeroes such a
 *     CC series (similarity_search.py:633-642); d_kurtosis_out (n_series floats or NULL) receives the values. */
int bpmf_detect_gated_dev(const float *d_series, size_t pitch, size_t n_series, const float *d_gaussian_sample,
                          float num_dev, size_t num_samples, size_t half_window_samp, size_t shift_samp,
                          size_t search_win, int mode, const float *d_thr_cap, int use_kurtosis,
                          float max_kurtosis, float *d_kurtosis_out, void *d_workspace, size_t workspace_bytes,
                          unsigned *d_count, int *d_series_idx, int *d_index, float *d_value,
                          float *d_thr_value, size_t capacity, void *stream);

/* ---------------------------------------------------------------- (2d) backprojection peak picking
 * The index part of Beamformer.find_detections (BPMF/template_search.py:603-610): utils._detect_peaks(maxbeam,
 * mpd=minimum_interevent_time) -- rising-edge local maxima, then suppression by decreasing height of every peak
 * closer than mpd samples to a kept one (BPMF/utils.py:2292-2345) -- followed by maxbeam[peak] > threshold[peak]
 * (threshold may be NULL: _detect_peaks alone).  Among exactly equal heights the later peak wins (the reference
 * leaves that order to NumPy's unstable default sort).  The host form returns the peaks in increasing order; the
 * device form writes them unsorted and d_count[0] = their number (> capacity means truncated). */
size_t bpmf_detect_peaks_workspace_bytes(size_t n_samples);
int bpmf_detect_peaks(const float *x, const float *threshold, size_t n_samples, size_t mpd, int *peaks,
                      size_t capacity, size_t *n_peaks);
int bpmf_detect_peaks_dev(const float *d_x, const float *d_threshold, size_t n_samples, size_t mpd,
                          void *d_workspace, size_t workspace_bytes, int *d_peaks, size_t capacity,
                          unsigned *d_count, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* BPMF_B200_H */
