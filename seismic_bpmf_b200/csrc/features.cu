// features.cu -- detection traces for the beamformer, built into libbpmf_b200_features.so.
//
// The step immediately upstream of beamforming in BPMF, all HBM-bound streaming work:
//   * envelope             BPMF/template_search.py:1599-1617  |scipy.signal.hilbert(trace)|: FFT of the whole
//                          record, one-sided spectrum doubled, inverse FFT, modulus.  The transforms are cuFFT
//                          (library code, like the reference's pocketfft); the spectrum weighting and the
//                          modulus are fused kernels here.
//   * saturated_envelopes  template_search.py:1525-1574  per trace: samples that are exactly 0 are "missing";
//                          more than half missing, or MAD < anomaly_threshold -> the trace is zeroed; else
//                          (x - median) / MAD over the present samples, missing kept at 0, clipped from above at
//                          max_dynamic_range.  Medians are order statistics of a radix sort (CUB), evaluated with
//                          the float32 arithmetic of np.median: given the same envelopes the result is bit-exact.
//   * time_dependent_threshold (backprojection flavour)  template_search.py:1418-1487  running median + CNR *
//                          running MAD over sliding windows, linearly interpolated to every sample.
// This library is separate from libbpmf_b200.so so that the matched-filter / beamforming library never depends
// on cuFFT.
#include <cub/cub.cuh>
#include <cufft.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>

#include <algorithm>
#include <mutex>

#include "../../include/bpmf_b200_features.h"
#include "common.cuh"

namespace bpmf {

// minimal error plumbing of this library (common.cuh declares them; abi.cu defines them for the other library)
static thread_local char g_feat_err[512];
void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_feat_err, sizeof(g_feat_err), fmt, ap);
    va_end(ap);
}
std::atomic<unsigned long long> g_launches{0};
MainKernelTimer::MainKernelTimer(cudaStream_t s) : st(s), on(false) {}
MainKernelTimer::~MainKernelTimer() {}

#define FEAT_CUFFT(expr)                                                                              \
    do {                                                                                              \
        cufftResult _r = (expr);                                                                      \
        if (_r != CUFFT_SUCCESS) {                                                                    \
            ::bpmf::set_error("%s:%d: %s -> cufft error %d", __FILE__, __LINE__, #expr, (int)_r);     \
            return BPMF_E_CUDA;                                                                       \
        }                                                                                             \
    } while (0)

// ---------------------------------------------------------------- envelope
// analytic-signal weights of scipy.signal.hilbert: h[0] = 1, h[N/2] = 1 (N even), 2 for 0 < k < N/2 (N even) or
// 0 < k <= (N-1)/2 (N odd), 0 above.  `spec` holds the N/2+1 bins of the R2C transform; `full` the N bins of Z.
__global__ void analytic_spectrum_kernel(const cufftComplex *__restrict__ spec, size_t N, cufftComplex *__restrict__ full) {
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= N) return;
    const size_t half = N / 2;
    cufftComplex z = make_cuFloatComplex(0.f, 0.f);
    if (k <= half) {
        float h;
        if (k == 0) h = 1.f;
        else if ((N & 1) == 0 && k == half) h = 1.f;
        else h = 2.f;
        const cufftComplex x = spec[k];
        z = make_cuFloatComplex(x.x * h, x.y * h);
    }
    full[k] = z;
}
__global__ void modulus_kernel(const cufftComplex *__restrict__ z, size_t N, float inv_n, float *__restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const float a = z[i].x * inv_n, b = z[i].y * inv_n;
    out[i] = hypotf(a, b);  // like np.abs(complex): no underflow of a*a for physical-unit amplitudes (~1e-20)
}

struct FftPlans {
    size_t n = 0;
    cufftHandle r2c = 0, c2c = 0;
    bool ok = false;
};
static std::mutex g_plan_mutex;
static FftPlans g_plans;

static int envelope_rows(const float *d_traces, size_t rows, size_t N, cufftComplex *d_spec, cufftComplex *d_full,
                         float *d_out, cudaStream_t st) {
    std::lock_guard<std::mutex> lock(g_plan_mutex);
    if (!g_plans.ok || g_plans.n != N) {
        if (g_plans.ok) {
            cufftDestroy(g_plans.r2c);
            cufftDestroy(g_plans.c2c);
            g_plans.ok = false;
        }
        BPMF_REQUIRE(N < (size_t)1 << 31, "envelope: record too long for a 32-bit cuFFT plan (%zu samples)", N);
        FEAT_CUFFT(cufftPlan1d(&g_plans.r2c, (int)N, CUFFT_R2C, 1));
        FEAT_CUFFT(cufftPlan1d(&g_plans.c2c, (int)N, CUFFT_C2C, 1));
        g_plans.n = N;
        g_plans.ok = true;
    }
    FEAT_CUFFT(cufftSetStream(g_plans.r2c, st));
    FEAT_CUFFT(cufftSetStream(g_plans.c2c, st));
    const unsigned blocks = (unsigned)((N + 255) / 256);
    for (size_t r = 0; r < rows; r++) {
        // cuFFT wants its real input aligned like a complex number: rows of odd length are staged in the
        // full-spectrum buffer first (free until the R2C transform has produced `spec`)
        const float *row = d_traces + r * N;
        if (((uintptr_t)row & 15u) != 0) {
            float *aligned = reinterpret_cast<float *>(d_full);
            BPMF_CUDA(cudaMemcpyAsync(aligned, row, sizeof(float) * N, cudaMemcpyDeviceToDevice, st));
            row = aligned;
        }
        FEAT_CUFFT(cufftExecR2C(g_plans.r2c, const_cast<float *>(row), d_spec));
        analytic_spectrum_kernel<<<blocks, 256, 0, st>>>(d_spec, N, d_full);
        BPMF_LAUNCH_CHECK();
        FEAT_CUFFT(cufftExecC2C(g_plans.c2c, d_full, d_full, CUFFT_INVERSE));
        modulus_kernel<<<blocks, 256, 0, st>>>(d_full, N, 1.0f / (float)N, d_out + r * N);
        BPMF_LAUNCH_CHECK();
    }
    return BPMF_OK;
}

// ---------------------------------------------------------------- order statistics helpers
// np.median of m sorted float32 values starting at p: middle element, or the float32 mean of the two middle ones
__device__ __forceinline__ float median_sorted(const float *p, size_t m) {
    if (m == 0) return nanf("");
    if (m & 1) return p[m / 2];
    return __fmul_rn(__fadd_rn(p[m / 2 - 1], p[m / 2]), 0.5f);
}

// per trace: count the missing (== 0) samples
__global__ void count_missing_kernel(const float *__restrict__ env, size_t N, unsigned long long *__restrict__ n_missing) {
    const size_t row = blockIdx.y;
    const float *x = env + row * N;
    unsigned long long c = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (size_t)gridDim.x * blockDim.x)
        c += (x[i] == 0.f) ? 1ull : 0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(n_missing + row, c);
}
// keys for the first sort: present samples keep their value, missing ones go to the end (+inf)
__global__ void present_keys_kernel(const float *__restrict__ env, size_t total, float *__restrict__ keys) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const float v = env[i];
    keys[i] = (v == 0.f) ? INFINITY : v;
}
__global__ void row_median_kernel(const float *__restrict__ sorted, size_t N, const unsigned long long *__restrict__ n_missing,
                                  size_t rows, float *__restrict__ med) {
    const size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const size_t m = N - (size_t)n_missing[r];
    med[r] = median_sorted(sorted + r * N, m);
}
// keys for the second sort: |x - median| of the present samples (float32, like np.abs(x - med)), missing -> +inf
__global__ void deviation_keys_kernel(const float *__restrict__ env, size_t N, const float *__restrict__ med,
                                      float *__restrict__ keys) {
    const size_t row = blockIdx.y;
    const float mu = med[row];
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (size_t)gridDim.x * blockDim.x) {
        const float v = env[row * N + i];
        keys[row * N + i] = (v == 0.f) ? INFINITY : fabsf(__fsub_rn(v, mu));
    }
}
// (x - median) / MAD, missing samples stay 0, clipped from above; traces with too many gaps or a tiny MAD are zeroed
__global__ void standardize_kernel(const float *__restrict__ env, size_t N, const unsigned long long *__restrict__ n_missing,
                                   const float *__restrict__ med, const float *__restrict__ mad, float anomaly_threshold,
                                   float max_dynamic_range, float *__restrict__ out, int *__restrict__ valid) {
    const size_t row = blockIdx.y;
    const bool too_many = (double)n_missing[row] > (double)N / 2.0;
    const float mu = med[row], sd = mad[row];
    const bool ok = !too_many && !(sd < anomaly_threshold);
    if (blockIdx.x == 0 && threadIdx.x == 0) valid[row] = ok ? 1 : 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (size_t)gridDim.x * blockDim.x) {
        const float v = env[row * N + i];
        float y = 0.f;
        if (ok && v != 0.f) {
            y = __fdiv_rn(__fsub_rn(v, mu), sd);
            y = (y > max_dynamic_range) ? max_dynamic_range : y;  // np.clip(x, x, max): NaN passes through
        }
        out[row * N + i] = y;
    }
}
__global__ void availability_kernel(const int *__restrict__ valid, int S, int C, int *__restrict__ avail) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    int a = 0;
    for (int c = 0; c < C; c++) a += valid[s * C + c];
    avail[s] = a;
}

struct RowOffsets {
    size_t N;
    __host__ __device__ size_t operator()(size_t r) const { return r * N; }
};

// sort every row of `keys` (rows x N) ascending into `sorted`
static int sort_rows(const float *d_keys, float *d_sorted, size_t rows, size_t N, void *d_temp, size_t temp_bytes,
                     cudaStream_t st) {
    for (size_t r = 0; r < rows; r++) {
        size_t need = temp_bytes;
        BPMF_CUDA(cub::DeviceRadixSort::SortKeys(d_temp, need, d_keys + r * N, d_sorted + r * N, N, 0, 32, st));
    }
    return BPMF_OK;
}
static size_t sort_temp_bytes(size_t N) {
    size_t need = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, need, (const float *)nullptr, (float *)nullptr, N, 0, 32, nullptr);
    return need;
}

static int standardize_rows(const float *d_env, size_t S, size_t C, size_t N, float anomaly_threshold,
                            float max_dynamic_range, float *d_out, int *d_avail, cudaStream_t st) {
    const size_t rows = S * C;
    DevBuf keys, sorted, temp, nmiss, med, mad, valid;
    BPMF_CUDA(keys.alloc(sizeof(float) * rows * N));
    BPMF_CUDA(sorted.alloc(sizeof(float) * rows * N));
    const size_t tb = sort_temp_bytes(N);
    BPMF_CUDA(temp.alloc(tb));
    BPMF_CUDA(nmiss.alloc(sizeof(unsigned long long) * rows));
    BPMF_CUDA(med.alloc(sizeof(float) * rows));
    BPMF_CUDA(mad.alloc(sizeof(float) * rows));
    BPMF_CUDA(valid.alloc(sizeof(int) * rows));
    BPMF_CUDA(cudaMemsetAsync(nmiss.p, 0, sizeof(unsigned long long) * rows, st));
    const dim3 grid2((unsigned)std::min<size_t>(1024, (N + 255) / 256), (unsigned)rows);
    count_missing_kernel<<<grid2, 256, 0, st>>>(d_env, N, nmiss.as<unsigned long long>());
    BPMF_LAUNCH_CHECK();
    present_keys_kernel<<<(unsigned)((rows * N + 255) / 256), 256, 0, st>>>(d_env, rows * N, keys.as<float>());
    BPMF_LAUNCH_CHECK();
    int rc = sort_rows(keys.as<float>(), sorted.as<float>(), rows, N, temp.p, tb, st);
    if (rc != BPMF_OK) return rc;
    row_median_kernel<<<(unsigned)((rows + 127) / 128), 128, 0, st>>>(sorted.as<float>(), N, nmiss.as<unsigned long long>(),
                                                                     rows, med.as<float>());
    BPMF_LAUNCH_CHECK();
    deviation_keys_kernel<<<grid2, 256, 0, st>>>(d_env, N, med.as<float>(), keys.as<float>());
    BPMF_LAUNCH_CHECK();
    rc = sort_rows(keys.as<float>(), sorted.as<float>(), rows, N, temp.p, tb, st);
    if (rc != BPMF_OK) return rc;
    row_median_kernel<<<(unsigned)((rows + 127) / 128), 128, 0, st>>>(sorted.as<float>(), N, nmiss.as<unsigned long long>(),
                                                                     rows, mad.as<float>());
    BPMF_LAUNCH_CHECK();
    standardize_kernel<<<grid2, 256, 0, st>>>(d_env, N, nmiss.as<unsigned long long>(), med.as<float>(), mad.as<float>(),
                                              anomaly_threshold, max_dynamic_range, d_out, valid.as<int>());
    BPMF_LAUNCH_CHECK();
    availability_kernel<<<(unsigned)((S + 127) / 128), 128, 0, st>>>(valid.as<int>(), (int)S, (int)C, d_avail);
    BPMF_LAUNCH_CHECK();
    BPMF_CUDA(cudaStreamSynchronize(st));  // the scratch buffers above are freed on return
    return BPMF_OK;
}

// ---------------------------------------------------------------- backprojection threshold
// window i (1-based, template_search.py:1453-1455): [i * shift, min(n, i * shift + window))
__global__ void gather_windows_kernel(const float *__restrict__ x, size_t n, size_t window, size_t shift, size_t n_win,
                                      const float *__restrict__ med, float *__restrict__ keys) {
    const size_t w = blockIdx.y;  // 0-based: window i = w + 1
    const size_t i1 = (w + 1) * shift;
    const size_t i2 = (i1 + window < n) ? i1 + window : n;
    for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < window; j += (size_t)gridDim.x * blockDim.x) {
        float v = INFINITY;  // beyond a truncated last window
        if (i1 + j < i2) {
            v = x[i1 + j];
            if (med) v = fabsf(__fsub_rn(v, med[w]));
        }
        keys[w * window + j] = v;
    }
}
__global__ void window_median_kernel(const float *__restrict__ sorted, size_t n, size_t window, size_t shift, size_t n_win,
                                     float *__restrict__ out) {
    const size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_win) return;
    const size_t i1 = (w + 1) * shift;
    const size_t i2 = (i1 + window < n) ? i1 + window : n;
    out[w] = median_sorted(sorted + w * window, i2 > i1 ? i2 - i1 : 0);
}
// knots (template_search.py:1466-1475): time[0] = 0, time[i] = (i1 + i2) / 2 (float32), time[-1] = n; values
// med + CNR * mad (float32), ends copy their neighbours; linear interpolation evaluated in float64
__global__ void interp_threshold_kernel(const float *__restrict__ med, const float *__restrict__ mad, size_t n, size_t window,
                                        size_t shift, size_t n_win, float cnr, double *__restrict__ thr) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    auto knot_time = [&](size_t k) -> double {  // k in [0, n_win + 1]
        if (k == 0) return 0.0;
        if (k == n_win + 1) return (double)(float)n;
        const size_t i1 = k * shift;
        const size_t i2 = (i1 + window < n) ? i1 + window : n;
        return (double)(float)((double)(i1 + i2) / 2.0);
    };
    auto knot_val = [&](size_t k) -> double {
        const size_t w = (k == 0) ? 0 : ((k == n_win + 1) ? n_win - 1 : k - 1);
        return (double)__fadd_rn(med[w], __fmul_rn(cnr, mad[w]));
    };
    const double x = (double)t;
    // last knot with time <= x (knots are non-decreasing)
    size_t lo = 0, hi = n_win + 1;
    while (hi - lo > 1) {
        const size_t mid = (lo + hi) / 2;
        if (knot_time(mid) <= x) lo = mid; else hi = mid;
    }
    const double x0 = knot_time(lo), x1 = knot_time(lo + 1), y0 = knot_val(lo), y1 = knot_val(lo + 1);
    thr[t] = (x1 > x0) ? y0 + (y1 - y0) * (x - x0) / (x1 - x0) : y0;
}

}  // namespace bpmf

using namespace bpmf;

extern "C" const char *bpmf_features_last_error(void) { return g_feat_err; }

extern "C" int bpmf_envelope(const float *traces, size_t n_traces, size_t n_samples, float *envelopes) {
    if (n_traces == 0 || n_samples == 0) return BPMF_OK;
    DevBuf in, out, spec, full;
    const size_t N = n_samples;
    BPMF_CUDA(in.alloc(sizeof(float) * n_traces * N));
    BPMF_CUDA(out.alloc(sizeof(float) * n_traces * N));
    BPMF_CUDA(spec.alloc(sizeof(cufftComplex) * (N / 2 + 1)));
    BPMF_CUDA(full.alloc(sizeof(cufftComplex) * N));
    cudaStream_t st = nullptr;
    BPMF_CUDA(cudaMemcpyAsync(in.p, traces, sizeof(float) * n_traces * N, cudaMemcpyHostToDevice, st));
    const int rc = envelope_rows(in.as<float>(), n_traces, N, spec.as<cufftComplex>(), full.as<cufftComplex>(),
                                 out.as<float>(), st);
    if (rc != BPMF_OK) return rc;
    BPMF_CUDA(cudaMemcpyAsync(envelopes, out.p, sizeof(float) * n_traces * N, cudaMemcpyDeviceToHost, st));
    BPMF_CUDA(cudaStreamSynchronize(st));
    return BPMF_OK;
}

extern "C" int bpmf_standardize_envelopes(const float *envelopes, size_t n_stations, size_t n_components,
                                          size_t n_samples, float anomaly_threshold, float max_dynamic_range,
                                          float *features, int *data_availability) {
    const size_t rows = n_stations * n_components, N = n_samples;
    if (rows == 0 || N == 0) return BPMF_OK;
    DevBuf env, out, avail;
    BPMF_CUDA(env.alloc(sizeof(float) * rows * N));
    BPMF_CUDA(out.alloc(sizeof(float) * rows * N));
    BPMF_CUDA(avail.alloc(sizeof(int) * n_stations));
    cudaStream_t st = nullptr;
    BPMF_CUDA(cudaMemcpyAsync(env.p, envelopes, sizeof(float) * rows * N, cudaMemcpyHostToDevice, st));
    const int rc = standardize_rows(env.as<float>(), n_stations, n_components, N, anomaly_threshold, max_dynamic_range,
                                    out.as<float>(), avail.as<int>(), st);
    if (rc != BPMF_OK) return rc;
    BPMF_CUDA(cudaMemcpyAsync(features, out.p, sizeof(float) * rows * N, cudaMemcpyDeviceToHost, st));
    BPMF_CUDA(cudaMemcpyAsync(data_availability, avail.p, sizeof(int) * n_stations, cudaMemcpyDeviceToHost, st));
    BPMF_CUDA(cudaStreamSynchronize(st));
    return BPMF_OK;
}

extern "C" int bpmf_saturated_envelopes(const float *traces, size_t n_stations, size_t n_components, size_t n_samples,
                                        float anomaly_threshold, float max_dynamic_range, float *features,
                                        int *data_availability) {
    const size_t rows = n_stations * n_components, N = n_samples;
    if (rows == 0 || N == 0) return BPMF_OK;
    DevBuf in, env, spec, full, avail;
    BPMF_CUDA(in.alloc(sizeof(float) * rows * N));
    BPMF_CUDA(env.alloc(sizeof(float) * rows * N));
    BPMF_CUDA(spec.alloc(sizeof(cufftComplex) * (N / 2 + 1)));
    BPMF_CUDA(full.alloc(sizeof(cufftComplex) * N));
    BPMF_CUDA(avail.alloc(sizeof(int) * n_stations));
    cudaStream_t st = nullptr;
    BPMF_CUDA(cudaMemcpyAsync(in.p, traces, sizeof(float) * rows * N, cudaMemcpyHostToDevice, st));
    int rc = envelope_rows(in.as<float>(), rows, N, spec.as<cufftComplex>(), full.as<cufftComplex>(), env.as<float>(), st);
    if (rc != BPMF_OK) return rc;
    rc = standardize_rows(env.as<float>(), n_stations, n_components, N, anomaly_threshold, max_dynamic_range,
                          in.as<float>(), avail.as<int>(), st);  // the input buffer is reused for the features
    if (rc != BPMF_OK) return rc;
    BPMF_CUDA(cudaMemcpyAsync(features, in.p, sizeof(float) * rows * N, cudaMemcpyDeviceToHost, st));
    BPMF_CUDA(cudaMemcpyAsync(data_availability, avail.p, sizeof(int) * n_stations, cudaMemcpyDeviceToHost, st));
    BPMF_CUDA(cudaStreamSynchronize(st));
    return BPMF_OK;
}

// Device-resident form: traces, features and data_availability are device pointers, work is enqueued on `stream`
// (the call synchronises that stream once before returning: its scratch buffers are freed on return).
// d_features may alias d_traces.
extern "C" int bpmf_saturated_envelopes_dev(const float *d_traces, size_t n_stations, size_t n_components,
                                            size_t n_samples, float anomaly_threshold, float max_dynamic_range,
                                            float *d_features, int *d_data_availability, void *stream) {
    const size_t rows = n_stations * n_components, N = n_samples;
    if (rows == 0 || N == 0) return BPMF_OK;
    DevBuf env, spec, full;
    BPMF_CUDA(env.alloc(sizeof(float) * rows * N));
    BPMF_CUDA(spec.alloc(sizeof(cufftComplex) * (N / 2 + 1)));
    BPMF_CUDA(full.alloc(sizeof(cufftComplex) * N));
    cudaStream_t st = (cudaStream_t)stream;
    int rc = envelope_rows(d_traces, rows, N, spec.as<cufftComplex>(), full.as<cufftComplex>(), env.as<float>(), st);
    if (rc != BPMF_OK) return rc;
    return standardize_rows(env.as<float>(), n_stations, n_components, N, anomaly_threshold, max_dynamic_range,
                            d_features, d_data_availability, st);
}

extern "C" int bpmf_bp_time_dependent_threshold(const float *network_response, size_t n_samples, size_t window,
                                                size_t shift, float cnr_threshold, double *threshold) {
    const size_t n = n_samples;
    if (n == 0) return BPMF_OK;
    BPMF_REQUIRE(window > 0 && shift > 0 && n >= window, "time_dependent_threshold: window %zu / shift %zu do not fit %zu samples",
                 window, shift, n);
    const size_t n_win = (n - window) / shift + 1;
    DevBuf x, keys, sorted, temp, med, mad, thr;
    BPMF_CUDA(x.alloc(sizeof(float) * n));
    BPMF_CUDA(keys.alloc(sizeof(float) * n_win * window));
    BPMF_CUDA(sorted.alloc(sizeof(float) * n_win * window));
    const size_t tb = sort_temp_bytes(window);
    BPMF_CUDA(temp.alloc(tb));
    BPMF_CUDA(med.alloc(sizeof(float) * n_win));
    BPMF_CUDA(mad.alloc(sizeof(float) * n_win));
    BPMF_CUDA(thr.alloc(sizeof(double) * n));
    cudaStream_t st = nullptr;
    BPMF_CUDA(cudaMemcpyAsync(x.p, network_response, sizeof(float) * n, cudaMemcpyHostToDevice, st));
    const dim3 grid((unsigned)std::min<size_t>(256, (window + 255) / 256), (unsigned)n_win);
    BPMF_REQUIRE(n_win <= 65535, "time_dependent_threshold: too many windows (%zu)", n_win);
    gather_windows_kernel<<<grid, 256, 0, st>>>(x.as<float>(), n, window, shift, n_win, nullptr, keys.as<float>());
    BPMF_LAUNCH_CHECK();
    int rc = sort_rows(keys.as<float>(), sorted.as<float>(), n_win, window, temp.p, tb, st);
    if (rc != BPMF_OK) return rc;
    window_median_kernel<<<(unsigned)((n_win + 127) / 128), 128, 0, st>>>(sorted.as<float>(), n, window, shift, n_win,
                                                                         med.as<float>());
    BPMF_LAUNCH_CHECK();
    gather_windows_kernel<<<grid, 256, 0, st>>>(x.as<float>(), n, window, shift, n_win, med.as<float>(), keys.as<float>());
    BPMF_LAUNCH_CHECK();
    rc = sort_rows(keys.as<float>(), sorted.as<float>(), n_win, window, temp.p, tb, st);
    if (rc != BPMF_OK) return rc;
    window_median_kernel<<<(unsigned)((n_win + 127) / 128), 128, 0, st>>>(sorted.as<float>(), n, window, shift, n_win,
                                                                         mad.as<float>());
    BPMF_LAUNCH_CHECK();
    interp_threshold_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(med.as<float>(), mad.as<float>(), n, window, shift,
                                                                        n_win, cnr_threshold, thr.as<double>());
    BPMF_LAUNCH_CHECK();
    BPMF_CUDA(cudaMemcpyAsync(threshold, thr.p, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
    BPMF_CUDA(cudaStreamSynchronize(st));
    return BPMF_OK;
}
