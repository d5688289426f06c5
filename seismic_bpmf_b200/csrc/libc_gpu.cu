// libc_gpu.cu -- CUDA implementations of the symbols BPMF/clib.py binds from BPMF/lib/libc.so
// (include/bpmf_b200.h group (1)).  Same names and signatures as BPMF/libc.c so the reference's
// clib.py loads this library unchanged; every function runs on the current CUDA device and
// reports failures on stderr (the reference ABI is `void`, BPMF/libc.h:1-11).
//
//   time_dependent_threshold  libc.c:516-673  intended (race-free) semantics, SURVEY section 5
//   select_cc_indexes         libc.c:441-485  closed form of the sequential scan (exact)
//   kurtosis                  libc.c:11-53    naive O(N*W), one thread per output
//   find_similar_moveouts[2]  libc.c:55-387   blocked greedy sweep, bit-exact distance arithmetic
#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "common.cuh"

namespace bpmf {

constexpr int GAUSS_LEN = 500;  // libc.c:526 ; clib.py:284

// ================================================================ time_dependent_threshold
struct ThrStat {
    double sum, devsum;
    unsigned long long count;
    float center, dev;
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// pass 0: sum + count of non-zero samples of the first n_head samples (libc.c:553-571)
// pass 1: sum of squared deviations from `center` (libc.c:574-586)
template <int PASS>
__global__ void thr_global_kernel(const float *__restrict__ x, size_t pitch, size_t n_head,
                                  ThrStat *__restrict__ stat) {
    const size_t b = blockIdx.y;
    const float *xs = x + b * pitch;
    const float center = (PASS == 1) ? stat[b].center : 0.f;
    double acc = 0.0;
    unsigned long long cnt = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_head;
         i += (size_t)gridDim.x * blockDim.x) {
        const float v = xs[i];
        if (v != 0.f) {
            if (PASS == 0) { acc += (double)v; cnt++; }
            else { const double d = (double)__fsub_rn(v, center); acc += d * d; }
        }
    }
    acc = warp_sum(acc);
    if (PASS == 0) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (PASS == 0) { atomicAdd(&stat[b].sum, acc); atomicAdd(&stat[b].count, cnt); }
        else atomicAdd(&stat[b].devsum, acc);
    }
}

__global__ void thr_finish_stat_kernel(ThrStat *stat, size_t B, int pass) {
    const size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const float cnt = (float)stat[b].count;
    if (pass == 0) stat[b].center = (float)stat[b].sum / cnt;          // libc.c:573
    else stat[b].dev = sqrtf((float)stat[b].devsum / cnt);             // libc.c:587
}

// replace zeros by center + gaussian[i % 500] * dev (libc.c:606-612), in place
__global__ void thr_fill_kernel(float *__restrict__ x, size_t pitch, size_t n,
                                const float *__restrict__ gauss, const ThrStat *__restrict__ stat) {
    const size_t b = blockIdx.y;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float *xs = x + b * pitch;
    if (xs[i] == 0.f) xs[i] = __fadd_rn(stat[b].center, __fmul_rn(gauss[i % GAUSS_LEN], stat[b].dev));
}

// sample i of a series as libc.c:606-612 leaves it: zeros replaced by center + gaussian[i % 500] * dev
__device__ __forceinline__ float thr_sample(const float *xs, size_t i, const float *gauss, float center, float dev) {
    const float v = xs[i];
    return (v == 0.f) ? __fadd_rn(center, __fmul_rn(gauss[i % GAUSS_LEN], dev)) : v;
}

// one CTA per (window, series): mean then rms deviation of W samples (libc.c:615-627); zeros are
// substituted on the fly, so the series itself need not have been filled
__global__ void thr_window_kernel(const float *__restrict__ x, size_t pitch, size_t W, size_t shift,
                                  size_t n_win, float num_dev, const float *__restrict__ gauss,
                                  const ThrStat *__restrict__ stat, float *__restrict__ tw) {
    __shared__ double red[32];
    __shared__ float mean_s;
    const size_t b = blockIdx.y, w = blockIdx.x;
    const float *xs = x + b * pitch;
    const size_t o = w * shift;
    const float center = stat[b].center, sdev = stat[b].dev;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    double acc = 0.0;
    for (size_t i = threadIdx.x; i < W; i += blockDim.x) acc += (double)thr_sample(xs, o + i, gauss, center, sdev);
    acc = warp_sum(acc);
    if (lane == 0) red[wid] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int q = 0; q < nw; q++) s += red[q];
        mean_s = (float)s / (float)W;  // _compute_average, libc.c:488-500
    }
    __syncthreads();
    const float mean = mean_s;
    acc = 0.0;
    for (size_t i = threadIdx.x; i < W; i += blockDim.x) {
        const double d = (double)__fsub_rn(thr_sample(xs, o + i, gauss, center, sdev), mean);
        acc += d * d;
    }
    acc = warp_sum(acc);
    __syncthreads();
    if (lane == 0) red[wid] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int q = 0; q < nw; q++) s += red[q];
        const float dev = sqrtf((float)s / (float)W);  // _compute_std, libc.c:502-514
        tw[b * n_win + w] = __fadd_rn(mean, __fmul_rn(num_dev, dev));  // libc.c:626
    }
}

// "delay the jump" smoothing (libc.c:631-651) evaluated in closed form per window, then the
// step-wise expansion (libc.c:654-669).  Interior index clamped to n_win-1.
__device__ __forceinline__ float thr_tw1(const float *tw, long long w) {
    if (w == 0) return tw[0];
    const float d0 = __fsub_rn(tw[w], tw[w - 1]);
    return (d0 < 0.f) ? __fsub_rn(tw[w], d0) : tw[w];
}

// threshold at sample i from the raw window values t[0..n_win)
__device__ __forceinline__ float thr_at(const float *t, size_t n_win, size_t n, size_t shift, size_t i) {
    long long w;
    if (i < shift) w = 0;
    else if (i >= n - shift) w = (long long)n_win - 1;
    else {
        w = (long long)(i / shift);
        if (w > (long long)n_win - 1) w = (long long)n_win - 1;
    }
    float v = thr_tw1(t, w);
    if (w < (long long)n_win - 1) {
        const float d1 = __fsub_rn(thr_tw1(t, w + 1), v);
        if (d1 > 0.f) v = __fadd_rn(v, d1);
    }
    return v;
}

__global__ void thr_expand_kernel(const float *__restrict__ tw, size_t n_win, size_t n, size_t shift,
                                  float *__restrict__ thr, size_t thr_pitch) {
    const size_t b = blockIdx.y;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    thr[b * thr_pitch + i] = thr_at(tw + b * n_win, n_win, n, shift, i);
}

// Excess kurtosis of each whole series, scipy.stats.kurtosis defaults (population moments, Fisher): the sanity check
// of MatchedFilter.find_detections (BPMF/similarity_search.py:633-642) zeroes a CC series whose kurtosis exceeds
// max_kurto.  Two passes in double precision: mean, then the 2nd and 4th central moments.
struct SeriesMoments {
    double sum, m2, m4;
    float kurtosis, pad;
};
template <int PASS>
__global__ void series_moments_kernel(const float *__restrict__ x, size_t pitch, size_t n,
                                      SeriesMoments *__restrict__ mo) {
    const size_t b = blockIdx.y;
    const float *xs = x + b * pitch;
    const double mean = (PASS == 1) ? mo[b].sum / (double)n : 0.0;
    double a2 = 0.0, a4 = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        if (PASS == 0) a2 += (double)xs[i];
        else {
            const double d = (double)xs[i] - mean, d2 = d * d;
            a2 += d2;
            a4 += d2 * d2;
        }
    }
    a2 = warp_sum(a2);
    if (PASS == 1) a4 = warp_sum(a4);
    if ((threadIdx.x & 31) == 0) {
        if (PASS == 0) atomicAdd(&mo[b].sum, a2);
        else { atomicAdd(&mo[b].m2, a2); atomicAdd(&mo[b].m4, a4); }
    }
}
__global__ void series_kurtosis_finish_kernel(SeriesMoments *mo, size_t B, size_t n, float *__restrict__ out) {
    const size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const double m2 = mo[b].m2 / (double)n, m4 = mo[b].m4 / (double)n;
    const float k = (m2 > 0.0) ? (float)(m4 / (m2 * m2) - 3.0) : nanf("");  // scipy returns nan for a constant series
    mo[b].kurtosis = k;
    if (out) out[b] = k;
}

// Candidate detections straight from the CC slab: sample i of series b is reported when
// cc > threshold(i) and, in mode 1, it also survives the declustering of select_cc_indexes
// (libc.c:441-485, closed form below).  Survivors are appended to a compact list; only that list has to
// leave the GPU.
__global__ void detect_kernel(const float *__restrict__ x, size_t pitch, const float *__restrict__ tw, size_t n_win,
                              size_t n, size_t shift, long long search_win, int mode, unsigned *__restrict__ count,
                              int *__restrict__ o_series, int *__restrict__ o_index, float *__restrict__ o_value,
                              float *__restrict__ o_thr, unsigned capacity, const float *__restrict__ thr_cap,
                              const SeriesMoments *__restrict__ mo, float max_kurtosis) {
    const size_t b = blockIdx.y;
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= (long long)n) return;
    // similarity_search.py:637-642: a series whose kurtosis exceeds max_kurto is zeroed -> it detects nothing
    if (mo != nullptr && mo[b].kurtosis > max_kurtosis) return;
    const float *cc = x + b * pitch;
    const float v = cc[j];
    float th = thr_at(tw + b * n_win, n_win, n, shift, (size_t)j);
    if (thr_cap != nullptr) th = fminf(thr_cap[b], th);  // similarity_search.py:629: np.minimum(max_CC_threshold * sum(w), thr)
    bool keep = v > th;
    if (keep && mode == 1) {
        const long long a = (j <= search_win) ? 0 : j - search_win;
        for (long long q = a; q < j; q++)
            if (cc[q] > v) { keep = false; break; }
        if (keep) {
            const long long e = (j + search_win < (long long)n - 1) ? j + search_win : (long long)n - 1;
            for (long long i = j + 1; i <= e; i++)
                if (!(cc[i] < v)) { keep = false; break; }
        }
    }
    if (keep) {
        const unsigned slot = atomicAdd(count, 1u);
        if (slot < capacity) {
            o_series[slot] = (int)b;
            o_index[slot] = (int)j;
            o_value[slot] = v;
            o_thr[slot] = th;
        }
    }
}

static bool thr_geometry(size_t n, size_t half, size_t shift, size_t *W, size_t *n_win) {
    *W = 2 * half;
    if (shift == 0 || *W == 0 || n < *W || *W < shift) return false;
    *n_win = (n - (*W - shift)) / shift;  // libc.c:528
    return *n_win >= 1;
}

}  // namespace bpmf

using namespace bpmf;

extern "C" size_t bpmf_threshold_workspace_bytes(size_t n_series, size_t n, size_t half, size_t shift) {
    size_t W, n_win;
    if (!thr_geometry(n, half, shift, &W, &n_win)) return 256;
    Workspace ws(nullptr, 0);
    ws.take<ThrStat>(n_series);
    ws.take<float>(n_series * n_win);
    ws.take<SeriesMoments>(n_series);
    return ws.used;
}

// global statistics + per-window mean/std -> raw window thresholds tw[B][n_win]; fills the series in place
// only when `fill` is set (the libc ABI mutates its input, libc.c:606-612)
static int thr_windows(float *d_series, size_t pitch, size_t B, const float *d_gauss, float num_dev, size_t n,
                       size_t W, size_t shift, size_t n_win, bool fill, ThrStat *stat, float *tw, cudaStream_t st) {
    BPMF_REQUIRE(B <= 65535, "time_dependent_threshold: more than 65535 series per call");
    BPMF_REQUIRE(n_win < ((size_t)1 << 31), "time_dependent_threshold: too many windows");
    BPMF_CUDA(cudaMemsetAsync(stat, 0, sizeof(ThrStat) * B, st));
    const size_t n_head = (n / W) * W;  // libc.c:553: only whole windows enter the global stats
    const unsigned gx = (unsigned)std::min<size_t>((n_head + 255) / 256, 1024);
    const unsigned nb = (unsigned)((n + 255) / 256);
    thr_global_kernel<0><<<dim3(gx ? gx : 1, (unsigned)B), 256, 0, st>>>(d_series, pitch, n_head, stat);
    BPMF_LAUNCH_CHECK();
    thr_finish_stat_kernel<<<(unsigned)((B + 127) / 128), 128, 0, st>>>(stat, B, 0);
    BPMF_LAUNCH_CHECK();
    thr_global_kernel<1><<<dim3(gx ? gx : 1, (unsigned)B), 256, 0, st>>>(d_series, pitch, n_head, stat);
    BPMF_LAUNCH_CHECK();
    thr_finish_stat_kernel<<<(unsigned)((B + 127) / 128), 128, 0, st>>>(stat, B, 1);
    BPMF_LAUNCH_CHECK();
    thr_window_kernel<<<dim3((unsigned)n_win, (unsigned)B), 256, 0, st>>>(d_series, pitch, W, shift, n_win, num_dev,
                                                                         d_gauss, stat, tw);
    BPMF_LAUNCH_CHECK();
    if (fill) {
        thr_fill_kernel<<<dim3(nb, (unsigned)B), 256, 0, st>>>(d_series, pitch, n, d_gauss, stat);
        BPMF_LAUNCH_CHECK();
    }
    return BPMF_OK;
}

extern "C" int bpmf_time_dependent_threshold_dev(float *d_series, size_t pitch, size_t B,
                                                 const float *d_gauss, float num_dev, size_t n,
                                                 size_t half, size_t shift, void *d_workspace,
                                                 size_t workspace_bytes, float *d_thr,
                                                 size_t thr_pitch, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (B == 0) return BPMF_OK;
    size_t W, n_win;
    BPMF_REQUIRE(thr_geometry(n, half, shift, &W, &n_win),
                 "time_dependent_threshold: invalid geometry n=%zu half_window=%zu shift=%zu", n, half, shift);
    Workspace ws(d_workspace, workspace_bytes);
    ThrStat *stat = ws.take<ThrStat>(B);
    float *tw = ws.take<float>(B * n_win);
    BPMF_REQUIRE(ws.ok(), "time_dependent_threshold: workspace too small");
    int rc = thr_windows(d_series, pitch, B, d_gauss, num_dev, n, W, shift, n_win, true, stat, tw, st);
    if (rc != BPMF_OK) return rc;
    thr_expand_kernel<<<dim3((unsigned)((n + 255) / 256), (unsigned)B), 256, 0, st>>>(tw, n_win, n, shift, d_thr,
                                                                                      thr_pitch);
    BPMF_LAUNCH_CHECK();
    return BPMF_OK;
}

extern "C" int bpmf_detect_gated_dev(const float *d_series, size_t pitch, size_t B, const float *d_gauss,
                                     float num_dev, size_t n, size_t half, size_t shift, size_t search_win, int mode,
                                     const float *d_thr_cap, int use_kurtosis, float max_kurtosis,
                                     float *d_kurtosis_out, void *d_workspace, size_t workspace_bytes,
                                     unsigned *d_count, int *d_series_idx, int *d_index, float *d_value,
                                     float *d_thr_value, size_t capacity, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    BPMF_CUDA(cudaMemsetAsync(d_count, 0, sizeof(unsigned), st));
    if (B == 0) return BPMF_OK;
    size_t W, n_win;
    BPMF_REQUIRE(thr_geometry(n, half, shift, &W, &n_win),
                 "detect: invalid geometry n=%zu half_window=%zu shift=%zu", n, half, shift);
    BPMF_REQUIRE(mode == 0 || mode == 1, "detect: mode must be 0 (above threshold) or 1 (select_cc_indexes)");
    BPMF_REQUIRE(capacity < ((size_t)1 << 32), "detect: capacity too large");
    Workspace ws(d_workspace, workspace_bytes);
    ThrStat *stat = ws.take<ThrStat>(B);
    float *tw = ws.take<float>(B * n_win);
    SeriesMoments *mo = ws.take<SeriesMoments>(B);
    BPMF_REQUIRE(ws.ok(), "detect: workspace too small");
    if (use_kurtosis) {
        BPMF_CUDA(cudaMemsetAsync(mo, 0, sizeof(SeriesMoments) * B, st));
        const dim3 grid(64, (unsigned)B);
        series_moments_kernel<0><<<grid, 256, 0, st>>>(d_series, pitch, n, mo);
        BPMF_LAUNCH_CHECK();
        series_moments_kernel<1><<<grid, 256, 0, st>>>(d_series, pitch, n, mo);
        BPMF_LAUNCH_CHECK();
        series_kurtosis_finish_kernel<<<(unsigned)((B + 127) / 128), 128, 0, st>>>(mo, B, n, d_kurtosis_out);
        BPMF_LAUNCH_CHECK();
    }
    // the series is NOT modified: zeros are substituted on the fly inside the window statistics
    int rc = thr_windows(const_cast<float *>(d_series), pitch, B, d_gauss, num_dev, n, W, shift, n_win, false, stat,
                         tw, st);
    if (rc != BPMF_OK) return rc;
    detect_kernel<<<dim3((unsigned)((n + 255) / 256), (unsigned)B), 256, 0, st>>>(
        d_series, pitch, tw, n_win, n, shift, (long long)search_win, mode, d_count, d_series_idx, d_index, d_value,
        d_thr_value, (unsigned)capacity, d_thr_cap, use_kurtosis ? mo : nullptr, max_kurtosis);
    BPMF_LAUNCH_CHECK();
    return BPMF_OK;
}

extern "C" int bpmf_detect_dev(const float *d_series, size_t pitch, size_t B, const float *d_gauss, float num_dev,
                               size_t n, size_t half, size_t shift, size_t search_win, int mode, void *d_workspace,
                               size_t workspace_bytes, unsigned *d_count, int *d_series_idx, int *d_index,
                               float *d_value, float *d_thr_value, size_t capacity, void *stream) {
    return bpmf_detect_gated_dev(d_series, pitch, B, d_gauss, num_dev, n, half, shift, search_win, mode, nullptr, 0,
                                 0.f, nullptr, d_workspace, workspace_bytes, d_count, d_series_idx, d_index, d_value,
                                 d_thr_value, capacity, stream);
}

static int thr_host(float *x, float *gauss, float num_dev, size_t n, size_t half, size_t shift, float *thr) {
    // the calling thread's own stream and scratch: called concurrently from a ThreadPoolExecutor
    // (similarity_search.py:590)
    const size_t wsb = bpmf_threshold_workspace_bytes(1, n, half, shift);
    HostPool &pool = host_pool();
    BPMF_CUDA(pool.begin(2 * sizeof(float) * n + sizeof(float) * GAUSS_LEN + wsb + 4 * 256));
    float *dx = pool.take<float>(n), *dt = pool.take<float>(n), *dg = pool.take<float>(GAUSS_LEN);
    char *dws = pool.take<char>(wsb);
    cudaStream_t st = pool.st;
    BPMF_CUDA(cudaMemcpyAsync(dx, x, sizeof(float) * n, cudaMemcpyHostToDevice, st));
    BPMF_CUDA(cudaMemcpyAsync(dg, gauss, sizeof(float) * GAUSS_LEN, cudaMemcpyHostToDevice, st));
    int rc = bpmf_time_dependent_threshold_dev(dx, n, 1, dg, num_dev, n, half, shift, dws, wsb, dt, n, st);
    if (rc != BPMF_OK) {
        cudaStreamSynchronize(st);
        return rc;
    }
    BPMF_CUDA(cudaMemcpyAsync(thr, dt, sizeof(float) * n, cudaMemcpyDeviceToHost, st));
    BPMF_CUDA(cudaMemcpyAsync(x, dx, sizeof(float) * n, cudaMemcpyDeviceToHost, st));  // zeros replaced, libc.c:606-612
    BPMF_CUDA(cudaStreamSynchronize(st));
    return BPMF_OK;
}

extern "C" void time_dependent_threshold(float *time_series, float *gaussian_sample, float num_dev,
                                         size_t num_samples, size_t half_window_samp,
                                         size_t shift_samp, int num_threads, float *threshold) {
    (void)num_threads;
    const int rc = thr_host(time_series, gaussian_sample, num_dev, num_samples, half_window_samp, shift_samp,
                            threshold);
    set_status(rc);
    if (rc != BPMF_OK) {
        fprintf(stderr, "libbpmf_b200: time_dependent_threshold failed: %s\n", bpmf_last_error());
        for (size_t i = 0; i < num_samples; i++) threshold[i] = NAN;  // `cc > NaN` is false: no false detections
    }
}

// ================================================================ select_cc_indexes
// The sequential scan of libc.c:455-484 has the closed form: index j is selected iff
//   cc[j] > threshold[j],  no cc[j'] > cc[j] for j' in [max(0, j-W), j),  and
//   cc[i] < cc[j] for every i in (j, min(j+W, n-1)]        (a later i with cc[i] >= cc[j] clears j).
__global__ void select_cc_kernel(const float *__restrict__ cc, const float *__restrict__ thr,
                                 long long W, long long n, int *__restrict__ sel) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const float v = cc[j];
    int keep = v > thr[j];
    if (keep) {
        const long long a = (j <= W) ? 0 : j - W;
        for (long long q = a; q < j; q++)
            if (cc[q] > v) { keep = 0; break; }
    }
    if (keep) {
        const long long b = (j + W < n - 1) ? j + W : n - 1;
        for (long long i = j + 1; i <= b; i++)
            if (!(cc[i] < v)) { keep = 0; break; }
    }
    sel[j] = keep;
}

static int select_host(float *cc, float *thr, size_t W, size_t n, int *sel) {
    if (n == 0) return BPMF_OK;
    HostPool &pool = host_pool();
    BPMF_CUDA(pool.begin(2 * sizeof(float) * n + sizeof(int) * n + 3 * 256));
    float *dc = pool.take<float>(n), *dt = pool.take<float>(n);
    int *ds = pool.take<int>(n);
    cudaStream_t st = pool.st;
    BPMF_CUDA(cudaMemcpyAsync(dc, cc, sizeof(float) * n, cudaMemcpyHostToDevice, st));
    BPMF_CUDA(cudaMemcpyAsync(dt, thr, sizeof(float) * n, cudaMemcpyHostToDevice, st));
    select_cc_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(dc, dt, (long long)W, (long long)n, ds);
    BPMF_LAUNCH_CHECK();
    BPMF_CUDA(cudaMemcpyAsync(sel, ds, sizeof(int) * n, cudaMemcpyDeviceToHost, st));
    BPMF_CUDA(cudaStreamSynchronize(st));
    return BPMF_OK;
}

extern "C" void select_cc_indexes(float *cc, float *threshold, size_t search_win, size_t n_corr,
                                  int *selection) {
    const int rc = select_host(cc, threshold, search_win, n_corr, selection);
    set_status(rc);
    if (rc != BPMF_OK) fprintf(stderr, "libbpmf_b200: select_cc_indexes failed: %s\n", bpmf_last_error());
}

// ================================================================ kurtosis
// libc.c:35-50: for n >= W, moments of signal[n-W : n]; written only when m2 > 1e-6.
__global__ void kurtosis_kernel(const float *__restrict__ sig, int W, long long length,
                                float *__restrict__ out) {
    extern __shared__ float tile[];
    const size_t ch = blockIdx.y;
    const long long n0 = (long long)blockIdx.x * blockDim.x + W;  // first output of this CTA
    const float *s = sig + ch * (size_t)length;
    const long long base = n0 - W;
    const int span = blockDim.x + W;
    for (int i = threadIdx.x; i < span; i += blockDim.x) {
        const long long g = base + i;
        tile[i] = (g < length) ? s[g] : 0.f;
    }
    __syncthreads();
    const long long n = n0 + threadIdx.x;
    if (n >= length) return;
    const float *w = tile + threadIdx.x;
    const double Wf = (double)W;
    double mean = 0.0;
    for (int i = 0; i < W; i++) mean += (double)w[i];
    mean /= Wf;
    double m2 = 0.0, m4 = 0.0;
    for (int i = 0; i < W; i++) {
        const double d = (double)w[i] - mean;
        const double d2 = d * d;
        m2 += d2;
        m4 += d2 * d2;
    }
    m2 /= Wf;
    m4 /= Wf;
    if ((float)m2 > 0.000001f)
        out[ch * (size_t)length + n] =
            (float)(1.0 / ((Wf - 2.0) * (Wf - 3.0)) * ((Wf * Wf - 1.0) * m4 / (m2 * m2) - 3.0 * (Wf - 1.0) * (Wf - 1.0)));
}

static int kurtosis_host(float *signal, int W, int S, int C, int length, float *kurto) {
    BPMF_REQUIRE(W >= 1 && S >= 0 && C >= 0 && length >= 0, "kurtosis: invalid shape");
    const size_t n_ch = (size_t)S * C, total = n_ch * (size_t)length;
    if (total == 0 || length <= W) return BPMF_OK;
    const size_t smem = sizeof(float) * (256 + (size_t)W);
    BPMF_REQUIRE(smem <= 200 * 1024, "kurtosis: window of %d samples exceeds the shared-memory tile", W);
    BPMF_REQUIRE(n_ch <= 65535, "kurtosis: too many channels");
    DevBuf ds, dk;
    BPMF_CUDA(ds.alloc(sizeof(float) * total));
    BPMF_CUDA(dk.alloc(sizeof(float) * total));
    BPMF_CUDA(cudaMemcpy(ds.p, signal, sizeof(float) * total, cudaMemcpyHostToDevice));
    // outputs the kernel does not write keep the caller's value (clib.py:91 passes zeros)
    BPMF_CUDA(cudaMemcpy(dk.p, kurto, sizeof(float) * total, cudaMemcpyHostToDevice));
    BPMF_CUDA(cudaFuncSetAttribute(kurtosis_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)(((size_t)(length - W) + 255) / 256), (unsigned)n_ch);
    kurtosis_kernel<<<grid, 256, smem>>>(ds.as<float>(), W, (long long)length, dk.as<float>());
    BPMF_LAUNCH_CHECK();
    BPMF_CUDA(cudaMemcpy(kurto, dk.p, sizeof(float) * total, cudaMemcpyDeviceToHost));
    return BPMF_OK;
}

extern "C" void kurtosis(float *signal, int W, int n_stations, int n_components, int length, float *kurto) {
    const int rc = kurtosis_host(signal, W, n_stations, n_components, length, kurto);
    set_status(rc);
    if (rc != BPMF_OK) {
        fprintf(stderr, "libbpmf_b200: kurtosis failed: %s\n", bpmf_last_error());
        const size_t n = (size_t)(n_stations > 0 ? n_stations : 0) * (size_t)(n_components > 0 ? n_components : 0) *
                         (size_t)(length > 0 ? length : 0);
        for (size_t i = 0; i < n; i++) kurto[i] = NAN;
    }
}

// ================================================================ find_similar_moveouts[2]
// The reference sweep (libc.c:121-170 per cell, :175-222 global; variant 2 at :288-334, :339-386)
// is greedy and order dependent: a non-redundant source n1 marks every later non-redundant n2 whose
// summed squared moveout difference is below n_for_diff * threshold^2.  It is evaluated here in
// blocks of FSM_B consecutive members: (A) one CTA resolves the block internally in index order,
// (B) the block's survivors are applied to every later member in parallel -- identical marks,
// because a source only ever acts as n1 after its own status is final.
namespace bpmf {
constexpr int FSM_B = 256;
constexpr int FSM_MAX_S = 128;

// dt2 += pow(a - b, 2) exactly as C evaluates it (float subtract, double square, float store)
__device__ __forceinline__ float acc_sq(float dt2, float a, float b) {
    const double d = (double)__fsub_rn(a, b);
    return (float)((double)dt2 + d * d);
}

// variant 2 ("closest"): stations ordered by n1's own moveouts (sort_indexes, libc.c:389-410)
__device__ float dist_closest(const float *__restrict__ m1, const float *__restrict__ m2,
                              const int *__restrict__ order, int n_for_diff) {
    float dt2 = 0.f;
    for (int s = 0; s < n_for_diff; s++) dt2 = acc_sq(dt2, m1[order[s]], m2[order[s]]);
    return dt2;
}

// variant 1 ("smallest"): the n_for_diff smallest squared differences, added in ascending order
__device__ float dist_smallest(const float *__restrict__ m1, const float *__restrict__ m2, int S,
                               int n_for_diff) {
    float sq[FSM_MAX_S];
    for (int s = 0; s < S; s++) {
        const double d = (double)__fsub_rn(m1[s], m2[s]);
        sq[s] = (float)(d * d);
    }
    float dt2 = 0.f;
    for (int i = 0; i < n_for_diff; i++) {  // partial selection sort (selectionSort, libc.c:421-439)
        int mi = i;
        for (int j = i + 1; j < S; j++)
            if (sq[j] < sq[mi]) mi = j;
        const float tmp = sq[mi]; sq[mi] = sq[i]; sq[i] = tmp;
        dt2 = __fadd_rn(dt2, sq[i]);
    }
    return dt2;
}

// selection argsort of each block member's moveouts (exactly libc.c:399-409, swaps included)
__global__ void fsm_order_kernel(const float *__restrict__ mv, const int *__restrict__ members,
                                 size_t blk0, size_t n_members, int S, int *__restrict__ order) {
    const size_t a = blk0 + threadIdx.x;
    if (threadIdx.x >= FSM_B || a >= n_members) return;
    const float *m = mv + (size_t)members[a] * S;
    int *o = order + (size_t)threadIdx.x * S;
    for (int i = 0; i < S; i++) o[i] = i;
    for (int i = 0; i + 1 < S; i++) {
        int mi = i;
        for (int j = i + 1; j < S; j++)
            if (m[o[j]] < m[o[mi]]) mi = j;
        const int tmp = o[mi]; o[mi] = o[i]; o[i] = tmp;
    }
}

template <int VARIANT>
__global__ void fsm_intra_kernel(const float *__restrict__ mv, const int *__restrict__ members,
                                 size_t blk0, size_t n_members, int S, int n_for_diff, float thr2,
                                 const int *__restrict__ order, int *__restrict__ red) {
    __shared__ int red_s[FSM_B];
    const int b = threadIdx.x;
    const size_t gb = blk0 + b;
    const bool live = gb < n_members;
    const int src_b = live ? members[gb] : 0;
    red_s[b] = live ? red[src_b] : 1;
    __syncthreads();
    const int nb = (int)((n_members - blk0 < (size_t)FSM_B) ? (n_members - blk0) : (size_t)FSM_B);
    for (int a = 0; a < nb; a++) {
        if (red_s[a] == 0 && b > a && live && red_s[b] == 0) {
            const float *m1 = mv + (size_t)members[blk0 + a] * S;
            const float *m2 = mv + (size_t)src_b * S;
            const float d = (VARIANT == 2) ? dist_closest(m1, m2, order + (size_t)a * S, n_for_diff)
                                           : dist_smallest(m1, m2, S, n_for_diff);
            if (d < thr2) red_s[b] = 1;
        }
        __syncthreads();
    }
    if (live) red[src_b] = red_s[b];
}

template <int VARIANT>
__global__ void fsm_apply_kernel(const float *__restrict__ mv, const int *__restrict__ members,
                                 size_t blk0, size_t n_members, int S, int n_for_diff, float thr2,
                                 const int *__restrict__ order, int *__restrict__ red) {
    __shared__ int surv[FSM_B];
    __shared__ int n_surv;
    const int nb = (int)((n_members - blk0 < (size_t)FSM_B) ? (n_members - blk0) : (size_t)FSM_B);
    if (threadIdx.x == 0) {
        int c = 0;
        for (int a = 0; a < nb; a++)
            if (red[members[blk0 + a]] == 0) surv[c++] = a;
        n_surv = c;
    }
    __syncthreads();
    const size_t g2 = blk0 + FSM_B + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g2 >= n_members) return;
    const int src2 = members[g2];
    if (red[src2]) return;
    const float *m2 = mv + (size_t)src2 * S;
    for (int q = 0; q < n_surv; q++) {
        const int a = surv[q];
        const float *m1 = mv + (size_t)members[blk0 + a] * S;
        const float d = (VARIANT == 2) ? dist_closest(m1, m2, order + (size_t)a * S, n_for_diff)
                                       : dist_smallest(m1, m2, S, n_for_diff);
        if (d < thr2) { red[src2] = 1; return; }
    }
}

template <int VARIANT>
static int fsm_sweep(const float *d_mv, const int *d_members, size_t n_members, int S, int n_for_diff,
                     float thr2, int *d_order, int *d_red) {
    for (size_t blk0 = 0; blk0 < n_members; blk0 += FSM_B) {
        if (VARIANT == 2) {
            fsm_order_kernel<<<1, FSM_B>>>(d_mv, d_members, blk0, n_members, S, d_order);
            BPMF_LAUNCH_CHECK();
        }
        fsm_intra_kernel<VARIANT><<<1, FSM_B>>>(d_mv, d_members, blk0, n_members, S, n_for_diff, thr2,
                                               d_order, d_red);
        BPMF_LAUNCH_CHECK();
        if (blk0 + FSM_B < n_members) {
            const size_t rest = n_members - (blk0 + FSM_B);
            fsm_apply_kernel<VARIANT><<<(unsigned)((rest + 127) / 128), 128>>>(
                d_mv, d_members, blk0, n_members, S, n_for_diff, thr2, d_order, d_red);
            BPMF_LAUNCH_CHECK();
        }
    }
    return BPMF_OK;
}

static int fsm_host(int variant, float *moveouts, float *lon, float *lat, float *cell_lon,
                    float *cell_lat, float threshold, size_t K, size_t S, size_t n_cl, size_t n_ct,
                    size_t n_for_diff, int *redundant) {
    if (K == 0) return BPMF_OK;
    BPMF_REQUIRE(S >= 1 && S <= FSM_MAX_S, "find_similar_moveouts: n_stations=%zu not in [1, %d]", S, FSM_MAX_S);
    BPMF_REQUIRE(n_for_diff <= S, "find_similar_moveouts: n_stations_for_diff > n_stations");
    BPMF_REQUIRE(K < ((size_t)1 << 31), "find_similar_moveouts: too many sources");
    const float thr2 = (float)((double)(float)n_for_diff * pow((double)threshold, 2.0));  // libc.c:79,249
    printf("RMS threshold: %.4f, summed squares threshold: %.4f\n", threshold, thr2);    // libc.c:80,250
    DevBuf d_mv, d_mem, d_order, d_red;
    BPMF_CUDA(d_mv.alloc(sizeof(float) * K * S));
    BPMF_CUDA(d_mem.alloc(sizeof(int) * K));
    BPMF_CUDA(d_order.alloc(sizeof(int) * FSM_B * S));
    BPMF_CUDA(d_red.alloc(sizeof(int) * K));
    BPMF_CUDA(cudaMemcpy(d_mv.p, moveouts, sizeof(float) * K * S, cudaMemcpyHostToDevice));
    BPMF_CUDA(cudaMemcpy(d_red.p, redundant, sizeof(int) * K, cudaMemcpyHostToDevice));
    std::vector<int> members;
    members.reserve(K);
    int rc;
    auto run = [&](void) -> int {
        if (members.size() < 2) return BPMF_OK;
        BPMF_CUDA(cudaMemcpy(d_mem.p, members.data(), sizeof(int) * members.size(), cudaMemcpyHostToDevice));
        return variant == 2 ? fsm_sweep<2>(d_mv.as<float>(), d_mem.as<int>(), members.size(), (int)S,
                                           (int)n_for_diff, thr2, d_order.as<int>(), d_red.as<int>())
                            : fsm_sweep<1>(d_mv.as<float>(), d_mem.as<int>(), members.size(), (int)S,
                                           (int)n_for_diff, thr2, d_order.as<int>(), d_red.as<int>());
    };
    // cell pre-pass: membership is index preparation on the host, the pair sweep runs on the device
    for (size_t i = 0; i < n_cl; i++)
        for (size_t j = 0; j < n_ct; j++) {
            members.clear();
            for (size_t k = 0; k < K; k++)
                if (!(lon[k] < cell_lon[i] || lon[k] >= cell_lon[i + 1] || lat[k] < cell_lat[j] ||
                      lat[k] >= cell_lat[j + 1]))
                    members.push_back((int)k);
            if ((rc = run()) != BPMF_OK) return rc;
        }
    members.resize(K);
    for (size_t k = 0; k < K; k++) members[k] = (int)k;
    if ((rc = run()) != BPMF_OK) return rc;
    BPMF_CUDA(cudaMemcpy(redundant, d_red.p, sizeof(int) * K, cudaMemcpyDeviceToHost));
    return BPMF_OK;
}
}  // namespace bpmf

extern "C" void find_similar_moveouts(float *moveouts, float *lon, float *lat, float *cell_lon,
                                      float *cell_lat, float threshold, size_t K, size_t S, size_t n_cl,
                                      size_t n_ct, size_t n_for_diff, int num_threads, int *redundant) {
    (void)num_threads;
    const int rc = fsm_host(1, moveouts, lon, lat, cell_lon, cell_lat, threshold, K, S, n_cl, n_ct, n_for_diff,
                            redundant);
    set_status(rc);
    if (rc != BPMF_OK) fprintf(stderr, "libbpmf_b200: find_similar_moveouts failed: %s\n", bpmf_last_error());
}

extern "C" void find_similar_moveouts2(float *moveouts, float *lon, float *lat, float *cell_lon,
                                       float *cell_lat, float threshold, size_t K, size_t S, size_t n_cl,
                                       size_t n_ct, size_t n_for_diff, int num_threads, int *redundant) {
    (void)num_threads;
    const int rc = fsm_host(2, moveouts, lon, lat, cell_lon, cell_lat, threshold, K, S, n_cl, n_ct, n_for_diff,
                            redundant);
    set_status(rc);
    if (rc != BPMF_OK) fprintf(stderr, "libbpmf_b200: find_similar_moveouts2 failed: %s\n", bpmf_last_error());
}

// ================================================================ helper symbols (libc.h:5-10)
extern "C" void swap(float *xp, float *yp) {
    const float t = *xp;
    *xp = *yp;
    *yp = t;
}

extern "C" void selectionSort(float *arr, int n) {
    for (int i = 0; i + 1 < n; i++) {
        int m = i;
        for (int j = i + 1; j < n; j++)
            if (arr[j] < arr[m]) m = j;
        swap(&arr[m], &arr[i]);
    }
}

extern "C" void sort_indexes(const float *moveouts, size_t *sorted_indexes, size_t n) {
    for (size_t i = 0; i < n; i++) sorted_indexes[i] = i;
    for (size_t i = 0; i + 1 < n; i++) {
        size_t m = i;
        for (size_t j = i + 1; j < n; j++)
            if (moveouts[sorted_indexes[j]] < moveouts[sorted_indexes[m]]) m = j;
        const size_t t = sorted_indexes[m];
        sorted_indexes[m] = sorted_indexes[i];
        sorted_indexes[i] = t;
    }
}

extern "C" float _compute_average(float *x, size_t len_x) {
    float a = 0.f;
    for (size_t i = 0; i < len_x; i++) a += x[i];
    return a / (float)len_x;
}

extern "C" float _compute_std(float *x, float x_mean, size_t len_x) {
    float ss = 0.f;
    for (size_t i = 0; i < len_x; i++) ss = (float)((double)ss + pow((double)(x[i] - x_mean), 2.0));
    return sqrtf(ss / (float)len_x);
}
