// common.cuh -- shared helpers for libbpmf_b200.so (sm_100a only)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/bpmf_b200.h"

namespace bpmf {

// ---------------------------------------------------------------- error plumbing
void set_error(const char *fmt, ...);
void set_status(int rc);  // outcome of the calling thread's last void-ABI call (bpmf_last_status)
extern std::atomic<unsigned long long> g_launches;
inline void count_launch(unsigned n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define BPMF_CUDA(expr)                                                                   \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess) {                                                          \
            ::bpmf::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,               \
                              cudaGetErrorString(_e));                                    \
            return BPMF_E_CUDA;                                                           \
        }                                                                                 \
    } while (0)

#define BPMF_LAUNCH_CHECK()                                                               \
    do {                                                                                  \
        ::bpmf::count_launch();                                                           \
        cudaError_t _e = cudaGetLastError();                                              \
        if (_e != cudaSuccess) {                                                          \
            ::bpmf::set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__,           \
                              cudaGetErrorString(_e));                                    \
            return BPMF_E_CUDA;                                                           \
        }                                                                                 \
    } while (0)

#define BPMF_REQUIRE(cond, ...)                                                           \
    do {                                                                                  \
        if (!(cond)) {                                                                    \
            ::bpmf::set_error(__VA_ARGS__);                                               \
            return BPMF_E_ARG;                                                            \
        }                                                                                 \
    } while (0)

// brackets the dominant kernel with events when bpmf_set_kernel_timing(1) is active
struct MainKernelTimer {
    cudaStream_t st;
    bool on;
    explicit MainKernelTimer(cudaStream_t s);
    ~MainKernelTimer();
};

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// bump allocator over the caller-provided device workspace (256-byte granules)
struct Workspace {
    char *base;
    size_t size, used;
    Workspace(void *p, size_t n) : base((char *)p), size(n), used(0) {}
    template <typename T>
    T *take(size_t count) {
        size_t bytes = align_up(count * sizeof(T), 256);
        T *r = (T *)(base ? base + used : nullptr);
        used += bytes;
        return r;
    }
    bool ok() const { return used <= size; }
};

// RAII device buffer for the host-pointer entry points
struct DevBuf {
    void *p = nullptr;
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
    ~DevBuf() {
        if (p) cudaFree(p);
    }
    template <typename T>
    T *as() const { return (T *)p; }
};

// Per-thread device scratch of the host-pointer libc entry points: one grow-only allocation and one stream per
// calling thread, reused across calls.  BPMF calls these symbols hundreds of times per day of data, from a thread
// pool (BPMF/similarity_search.py:590-593); cudaMalloc / cudaFree / stream creation per call cost more than the
// kernels themselves (and cudaFree synchronises the whole device, serialising the pool's threads).
struct HostPool {
    cudaStream_t st = nullptr;
    char *base = nullptr;
    size_t cap = 0, used = 0;
    int dev = -1;
    cudaError_t begin(size_t total_bytes) {
        int d = 0;
        cudaError_t e = cudaGetDevice(&d);
        if (e != cudaSuccess) return e;
        if (d != dev) { release(); dev = d; }
        if (!st && (e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking)) != cudaSuccess) return e;
        total_bytes += 4096;
        if (cap < total_bytes) {
            if (base) cudaFree(base);
            base = nullptr;
            cap = 0;
            if ((e = cudaMalloc((void **)&base, total_bytes + total_bytes / 4)) != cudaSuccess) return e;
            cap = total_bytes + total_bytes / 4;
        }
        used = 0;
        return cudaSuccess;
    }
    template <typename T>
    T *take(size_t count) {
        T *r = (T *)(base + used);
        used += align_up(count * sizeof(T), 256);
        return r;
    }
    void release() {
        if (base) cudaFree(base);
        if (st) cudaStreamDestroy(st);
        base = nullptr;
        st = nullptr;
        cap = used = 0;
    }
    // no destructor on purpose: at thread / process exit the CUDA context may already be gone
};
HostPool &host_pool();  // the calling thread's pool

int num_sms();
int reserved_sms();  // SMs the persistent kernels leave free (bpmf_set_reserved_sms)

// ---------------------------------------------------------------- matched filter internals
constexpr float MF_GUARD = 1e-6f;  // sum T^2 * sum D^2 <= guard -> exactly 0

struct MfShape {
    size_t T, S, C, L, N, step, n_corr;
    int normalize, network_sum;
    size_t SC() const { return S * C; }
};

// per-template valid output range [lo, hi] (inclusive; hi < lo -> nothing to compute)
struct MfMeta {
    float *sst;   // (T*SC) sum of squares of each template channel
    float *sumt;  // (T*SC) sum of each template channel          (normalize == full)
    int *lo, *hi; // (T)
};

int mf_prepass(const MfShape &sh, const float *d_templates, const int *d_moveouts,
               const float *d_weights, const float *d_data, MfMeta meta, float *d_energy,
               float *d_mean, cudaStream_t st);
// sliding window energy (or its rsqrt) per channel, row pitch `out_pitch` elements; phase8 writes
// out[ch][j % 8][j / 8] (rows of out_pitch) instead of out[ch][j]
int mf_energy(const MfShape &sh, const float *d_data, float *d_energy, float *d_mean, size_t out_pitch,
              bool rsqrt_mode, bool phase8, cudaStream_t st);
int mf_fp32_run(const MfShape &sh, const float *d_templates, const int *d_moveouts,
                const float *d_weights, const float *d_data, MfMeta meta, const float *d_energy,
                const float *d_mean, float *d_cc, size_t cc_pitch, cudaStream_t st);

// tensor-core (tcgen05) path: trace staging is prepared once and shared by every template batch
bool mf_tc_supported(size_t L, size_t step, int normalize, int network_sum);
size_t mf_tc_trace_ws_bytes(const MfShape &sh, bool want_lo);
size_t mf_tc_template_ws_bytes(const MfShape &sh, int split);
int mf_tc_prepare_traces(const MfShape &sh, bool want_lo, const float *d_data, Workspace &trace_ws,
                         cudaStream_t st);
int mf_tc_run(const MfShape &sh, int split, const float *d_templates, const int *d_moveouts,
              const float *d_weights, MfMeta meta, Workspace &trace_ws, bool traces_have_lo, Workspace &tpl_ws,
              float *d_cc, size_t cc_pitch, bool zero_cc, cudaStream_t st);

}  // namespace bpmf
