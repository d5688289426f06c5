// peaks.cu -- peak picking of the backprojection's composite network response on the device
// (SURVEY section 8 f2): the index part of Beamformer.find_detections, BPMF/template_search.py:603-610, i.e.
// utils._detect_peaks(maxbeam, mpd) (BPMF/utils.py:2292-2345, defaults: rising edges) followed by
// maxbeam[peak] > detection_threshold[peak].
//
// _detect_peaks suppresses peaks closer than mpd by DECREASING height: a sequential loop over the height-sorted
// candidates in the reference.  The same set is the fixed point of a parallel rule evaluated here in rounds:
//   an undecided candidate is KEPT when no undecided candidate within +-mpd has a higher priority, and
//   DELETED when a kept candidate lies within +-mpd
// (a kept peak only ever deletes lower-priority ones: a higher-priority undecided neighbour would have blocked it).
// Priority = (height, index): among exactly equal heights the later peak wins -- the order of
// ind[np.argsort(x[ind], kind="stable")][::-1]; the reference's default (unstable) sort leaves that order to the
// platform (oracle/peaks_numpy.py).
#include <math.h>

#include <algorithm>
#include <vector>

#include "common.cuh"

namespace bpmf {

enum : unsigned char { PK_NONE = 0, PK_UNDECIDED = 1, PK_KEPT = 2, PK_DELETED = 3 };

// rising-edge local maxima: x[i] > x[i-1] and x[i+1] <= x[i]; first / last sample and anything next to a NaN never
__global__ void peaks_mark_kernel(const float *__restrict__ x, long long n, unsigned char *__restrict__ state,
                                  int suppress) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned char s = PK_NONE;
    if (i >= 1 && i + 1 < n) {
        const float a = x[i - 1], b = x[i], c = x[i + 1];
        if (!isnan(a) && !isnan(b) && !isnan(c) && b > a && c <= b) s = suppress ? PK_UNDECIDED : PK_KEPT;
    }
    state[i] = s;
}

__global__ void peaks_round_kernel(const float *__restrict__ x, long long n, long long mpd,
                                   const unsigned char *__restrict__ in, unsigned char *__restrict__ out,
                                   unsigned *__restrict__ undecided_left) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned char s = in[i];
    if (s == PK_UNDECIDED) {
        const float v = x[i];
        const long long a = (i > mpd) ? i - mpd : 0, b = (i + mpd < n - 1) ? i + mpd : n - 1;
        bool blocked = false, killed = false;
        for (long long j = a; j <= b; j++) {
            const unsigned char sj = in[j];
            if (sj == PK_KEPT) { killed = true; break; }
            if (sj == PK_UNDECIDED && j != i) {
                const float u = x[j];
                if (u > v || (u == v && j > i)) blocked = true;
            }
        }
        s = killed ? PK_DELETED : (blocked ? PK_UNDECIDED : PK_KEPT);
        if (s == PK_UNDECIDED) atomicAdd(undecided_left, 1u);
    }
    out[i] = s;
}

__global__ void peaks_collect_kernel(const float *__restrict__ x, const float *__restrict__ thr, long long n,
                                     const unsigned char *__restrict__ state, int *__restrict__ peaks,
                                     unsigned capacity, unsigned *__restrict__ count) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || state[i] != PK_KEPT) return;
    if (thr != nullptr && !(x[i] > thr[i])) return;
    const unsigned k = atomicAdd(count, 1u);
    if (k < capacity) peaks[k] = (int)i;
}

}  // namespace bpmf

using namespace bpmf;

extern "C" size_t bpmf_detect_peaks_workspace_bytes(size_t n) {
    Workspace ws(nullptr, 0);
    ws.take<unsigned char>(n);
    ws.take<unsigned char>(n);
    ws.take<unsigned>(64);
    return ws.used;
}

// d_count[0] receives the number of peaks found (may exceed capacity: then only `capacity` were written); the peaks
// are NOT sorted.  Synchronises the stream once per suppression round (4 bytes read back).
extern "C" int bpmf_detect_peaks_dev(const float *d_x, const float *d_threshold, size_t n, size_t mpd, void *d_workspace,
                                     size_t workspace_bytes, int *d_peaks, size_t capacity, unsigned *d_count,
                                     void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    BPMF_CUDA(cudaMemsetAsync(d_count, 0, sizeof(unsigned), st));
    if (n < 3) return BPMF_OK;
    BPMF_REQUIRE(n < ((size_t)1 << 31), "detect_peaks: series too long for int32 indexes");
    BPMF_REQUIRE(capacity < ((size_t)1 << 32), "detect_peaks: capacity too large");
    Workspace ws(d_workspace, workspace_bytes);
    unsigned char *s0 = ws.take<unsigned char>(n), *s1 = ws.take<unsigned char>(n);
    unsigned *left = ws.take<unsigned>(64);
    BPMF_REQUIRE(ws.ok(), "detect_peaks: workspace too small (%zu < %zu)", workspace_bytes, ws.used);
    const unsigned grid = (unsigned)((n + 255) / 256);
    const int suppress = mpd > 1 ? 1 : 0;  // utils.py:2334: `if ind.size and mpd > 1`
    peaks_mark_kernel<<<grid, 256, 0, st>>>(d_x, (long long)n, s0, suppress);
    BPMF_LAUNCH_CHECK();
    if (suppress) {
        for (int round = 0; round < 100000; round++) {
            BPMF_CUDA(cudaMemsetAsync(left, 0, sizeof(unsigned), st));
            peaks_round_kernel<<<grid, 256, 0, st>>>(d_x, (long long)n, (long long)mpd, s0, s1, left);
            BPMF_LAUNCH_CHECK();
            std::swap(s0, s1);
            unsigned h_left = 0;
            BPMF_CUDA(cudaMemcpyAsync(&h_left, left, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
            BPMF_CUDA(cudaStreamSynchronize(st));
            if (h_left == 0) break;
        }
    }
    peaks_collect_kernel<<<grid, 256, 0, st>>>(d_x, d_threshold, (long long)n, s0, d_peaks, (unsigned)capacity, d_count);
    BPMF_LAUNCH_CHECK();
    return BPMF_OK;
}

// Host arrays in / out.  threshold may be NULL (then every suppressed-survivor peak is returned, _detect_peaks alone).
// peaks (capacity ints) receives the indexes in increasing order, *n_peaks their number.
extern "C" int bpmf_detect_peaks(const float *x, const float *threshold, size_t n, size_t mpd, int *peaks,
                                 size_t capacity, size_t *n_peaks) {
    *n_peaks = 0;
    if (n < 3) return BPMF_OK;
    const size_t wsb = bpmf_detect_peaks_workspace_bytes(n);
    HostPool &pool = host_pool();
    const size_t cap_dev = std::min(capacity, n);
    BPMF_CUDA(pool.begin(sizeof(float) * 2 * n + wsb + sizeof(int) * cap_dev + 5 * 256));
    float *dx = pool.take<float>(n), *dt = threshold ? pool.take<float>(n) : nullptr;
    char *dws = pool.take<char>(wsb);
    int *dp = pool.take<int>(cap_dev);
    unsigned *dc = pool.take<unsigned>(1);
    cudaStream_t st = pool.st;
    BPMF_CUDA(cudaMemcpyAsync(dx, x, sizeof(float) * n, cudaMemcpyHostToDevice, st));
    if (threshold) BPMF_CUDA(cudaMemcpyAsync(dt, threshold, sizeof(float) * n, cudaMemcpyHostToDevice, st));
    int rc = bpmf_detect_peaks_dev(dx, dt, n, mpd, dws, wsb, dp, cap_dev, dc, st);
    if (rc != BPMF_OK) {
        cudaStreamSynchronize(st);
        return rc;
    }
    unsigned cnt = 0;
    BPMF_CUDA(cudaMemcpyAsync(&cnt, dc, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    BPMF_CUDA(cudaStreamSynchronize(st));
    BPMF_REQUIRE(cnt <= cap_dev, "detect_peaks: %u peaks exceed the capacity %zu", cnt, capacity);
    BPMF_CUDA(cudaMemcpyAsync(peaks, dp, sizeof(int) * cnt, cudaMemcpyDeviceToHost, st));
    BPMF_CUDA(cudaStreamSynchronize(st));
    std::sort(peaks, peaks + cnt);
    *n_peaks = cnt;
    return BPMF_OK;
}
