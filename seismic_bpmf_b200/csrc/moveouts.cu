// moveouts.cu -- travel times (seconds) -> integer moveouts (samples), SURVEY.md 8f-f2.
//
// Replaces, for the (n_sources, n_stations, n_phases) table the beamformer consumes,
//   * BPMF/utils.py:1257-1270  sec_to_samp:  sign(t) * int64(|t * sr| + epsilon), evaluated in the dtype of t
//     (float32 tables stay float32: the rounding of t * sr decides which integer comes out);
//   * BPMF/template_search.py:207-219  get_travel_times_array(units="samples", relative_to_first=True):
//     subtract the per-source minimum over (station, phase); moveout_to_tt = minimum / sampling_rate.
// The reference rebuilds this table with Python loops on every access of Beamformer.moveouts
// (template_search.py:444-454), once per detection (template_search.py:636); here it is one HBM-bound pass
// (reads 4 or 8 bytes, writes 4 bytes per entry) whose int32 result stays on the device for bpmf_beamform_*_dev.
#include <limits.h>

#include <algorithm>

#include "common.cuh"

namespace bpmf {

__device__ __forceinline__ long long sec_to_samp_one(float t, float sr, float eps) {
    // no fused multiply-add: numpy rounds the product before adding epsilon
    const float x = __fadd_rn(fabsf(__fmul_rn(t, sr)), eps);
    const long long v = (long long)x;  // truncation, like np.int64()
    return (t > 0.f) ? v : ((t < 0.f) ? -v : 0ll);
}
__device__ __forceinline__ long long sec_to_samp_one(double t, double sr, double eps) {
    const double x = __dadd_rn(fabs(__dmul_rn(t, sr)), eps);
    const long long v = (long long)x;
    return (t > 0.0) ? v : ((t < 0.0) ? -v : 0ll);
}

// one warp per source: lanes stride over the (station, phase) entries
template <typename F>
__global__ void moveouts_kernel(const F *__restrict__ tt, size_t K, int SP, F sr, F eps, int relative,
                                int *__restrict__ mv, double *__restrict__ mv_to_tt, int *__restrict__ overflow) {
    const size_t k = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (k >= K) return;
    const F *row = tt + k * (size_t)SP;
    long long mn = LLONG_MAX;
    for (int i = lane; i < SP; i += 32) {
        const long long v = sec_to_samp_one(row[i], sr, eps);
        mn = v < mn ? v : mn;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const long long x = __shfl_xor_sync(0xffffffffu, mn, o);
        mn = x < mn ? x : mn;
    }
    const long long base = relative ? mn : 0ll;
    bool bad = false;
    for (int i = lane; i < SP; i += 32) {
        const long long v = sec_to_samp_one(row[i], sr, eps) - base;
        bad |= (v > 2147483647ll || v < -2147483648ll);
        mv[k * (size_t)SP + i] = (int)v;
    }
    if (bad) atomicExch(overflow, 1);
    if (lane == 0 && mv_to_tt) mv_to_tt[k] = relative ? (double)mn / (double)sr : 0.0;
}

}  // namespace bpmf

using namespace bpmf;

extern "C" int bpmf_moveouts_from_travel_times_dev(const void *d_travel_times, int is_float64, size_t n_sources,
                                                   size_t n_stations, size_t n_phases, double sampling_rate,
                                                   double epsilon, int relative_to_first, int *d_moveouts,
                                                   double *d_moveout_to_tt, int *d_overflow, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const size_t SP = n_stations * n_phases;
    if (n_sources == 0 || SP == 0) return BPMF_OK;
    BPMF_REQUIRE(SP < (1u << 30), "moveouts: too many stations x phases");
    BPMF_CUDA(cudaMemsetAsync(d_overflow, 0, sizeof(int), st));
    const unsigned blocks = (unsigned)((n_sources + 7) / 8);
    if (is_float64)
        moveouts_kernel<double><<<blocks, 256, 0, st>>>((const double *)d_travel_times, n_sources, (int)SP,
                                                        sampling_rate, epsilon, relative_to_first, d_moveouts,
                                                        d_moveout_to_tt, d_overflow);
    else
        moveouts_kernel<float><<<blocks, 256, 0, st>>>((const float *)d_travel_times, n_sources, (int)SP,
                                                       (float)sampling_rate, (float)epsilon, relative_to_first,
                                                       d_moveouts, d_moveout_to_tt, d_overflow);
    BPMF_LAUNCH_CHECK();
    return BPMF_OK;
}

extern "C" int bpmf_moveouts_from_travel_times(const void *travel_times, int is_float64, size_t n_sources,
                                               size_t n_stations, size_t n_phases, double sampling_rate,
                                               double epsilon, int relative_to_first, int *moveouts,
                                               double *moveout_to_tt) {
    const size_t n = n_sources * n_stations * n_phases;
    if (n == 0) return BPMF_OK;
    const size_t esz = is_float64 ? 8 : 4;
    DevBuf tt, mv, m2t, ovf;
    BPMF_CUDA(tt.alloc(esz * n));
    BPMF_CUDA(mv.alloc(sizeof(int) * n));
    BPMF_CUDA(m2t.alloc(sizeof(double) * n_sources));
    BPMF_CUDA(ovf.alloc(sizeof(int)));
    cudaStream_t st = nullptr;
    BPMF_CUDA(cudaMemcpyAsync(tt.p, travel_times, esz * n, cudaMemcpyHostToDevice, st));
    const int rc = bpmf_moveouts_from_travel_times_dev(tt.p, is_float64, n_sources, n_stations, n_phases,
                                                       sampling_rate, epsilon, relative_to_first, mv.as<int>(),
                                                       m2t.as<double>(), ovf.as<int>(), st);
    if (rc != BPMF_OK) return rc;
    int overflow = 0;
    BPMF_CUDA(cudaMemcpyAsync(moveouts, mv.p, sizeof(int) * n, cudaMemcpyDeviceToHost, st));
    if (moveout_to_tt)
        BPMF_CUDA(cudaMemcpyAsync(moveout_to_tt, m2t.p, sizeof(double) * n_sources, cudaMemcpyDeviceToHost, st));
    BPMF_CUDA(cudaMemcpyAsync(&overflow, ovf.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    BPMF_CUDA(cudaStreamSynchronize(st));
    BPMF_REQUIRE(!overflow, "moveouts: a travel time does not fit 32-bit samples at %g Hz", sampling_rate);
    return BPMF_OK;
}

// ---------------------------------------------------------------- trace normalisation
// MatchedFilter.set_data, BPMF/similarity_search.py:181-185: every (station, component) trace is divided in place by
// its standard deviation (population std, 0 -> 1) "to avoid problems with low numbers"; the norms are kept.
// Three HBM-bound passes: mean, squared deviations (float64 sums, one CTA row per channel), then the scaling.
namespace bpmf {
// pass 0: sum of the channel; pass 1: sum of squared deviations from the mean of pass 0 (two passes: E[x^2] - mean^2
// would lose the variance of a trace that rides on a large offset)
__global__ void trace_moments_kernel(const float *__restrict__ data, size_t N, int pass, double *__restrict__ sums) {
    const size_t ch = blockIdx.y;
    const float *d = data + ch * N;
    const double mean = pass ? sums[2 * ch] / (double)N : 0.0;
    double s = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (size_t)gridDim.x * blockDim.x) {
        const double v = (double)d[i] - mean;
        s += pass ? v * v : v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(sums + 2 * ch + pass, s);
}
__global__ void trace_scale_kernel(float *__restrict__ data, size_t N, const double *__restrict__ sums,
                                   float *__restrict__ norm) {
    const size_t ch = blockIdx.y;
    float sd = (float)sqrt(sums[2 * ch + 1] / (double)N);
    if (sd == 0.f) sd = 1.f;
    if (blockIdx.x == 0 && threadIdx.x == 0) norm[ch] = sd;
    float *d = data + ch * N;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (size_t)gridDim.x * blockDim.x)
        d[i] = __fdiv_rn(d[i], sd);
}
}  // namespace bpmf

extern "C" int bpmf_normalize_traces_dev(float *d_data, size_t n_channels, size_t n_samples, float *d_norm,
                                         double *d_scratch, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (n_channels == 0 || n_samples == 0) return BPMF_OK;
    BPMF_REQUIRE(n_channels <= 65535, "normalize_traces: too many channels");
    BPMF_CUDA(cudaMemsetAsync(d_scratch, 0, sizeof(double) * 2 * n_channels, st));
    const dim3 grid((unsigned)std::min<size_t>(256, (n_samples + 1023) / 1024), (unsigned)n_channels);
    trace_moments_kernel<<<grid, 256, 0, st>>>(d_data, n_samples, 0, d_scratch);
    BPMF_LAUNCH_CHECK();
    trace_moments_kernel<<<grid, 256, 0, st>>>(d_data, n_samples, 1, d_scratch);
    BPMF_LAUNCH_CHECK();
    trace_scale_kernel<<<grid, 256, 0, st>>>(d_data, n_samples, d_scratch, d_norm);
    BPMF_LAUNCH_CHECK();
    return BPMF_OK;
}

extern "C" int bpmf_normalize_traces(float *data, size_t n_channels, size_t n_samples, float *norm) {
    if (n_channels == 0 || n_samples == 0) return BPMF_OK;
    DevBuf d, nr, sc;
    BPMF_CUDA(d.alloc(sizeof(float) * n_channels * n_samples));
    BPMF_CUDA(nr.alloc(sizeof(float) * n_channels));
    BPMF_CUDA(sc.alloc(sizeof(double) * 2 * n_channels));
    cudaStream_t st = nullptr;
    BPMF_CUDA(cudaMemcpyAsync(d.p, data, sizeof(float) * n_channels * n_samples, cudaMemcpyHostToDevice, st));
    const int rc = bpmf_normalize_traces_dev(d.as<float>(), n_channels, n_samples, nr.as<float>(), sc.as<double>(), st);
    if (rc != BPMF_OK) return rc;
    BPMF_CUDA(cudaMemcpyAsync(data, d.p, sizeof(float) * n_channels * n_samples, cudaMemcpyDeviceToHost, st));
    BPMF_CUDA(cudaMemcpyAsync(norm, nr.p, sizeof(float) * n_channels, cudaMemcpyDeviceToHost, st));
    BPMF_CUDA(cudaStreamSynchronize(st));
    return BPMF_OK;
}
