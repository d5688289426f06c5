// abi.cu -- extern "C" surface of libbpmf_b200.so: error plumbing, device query and the
// matched-filter entry points (include/bpmf_b200.h groups (2a) and (3)).
#include <stdarg.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <vector>

#include <stdlib.h>

#include "common.cuh"

namespace bpmf {

static thread_local char g_err[512] = "";
std::atomic<unsigned long long> g_launches{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static thread_local HostPool g_pool;
HostPool &host_pool() { return g_pool; }

static thread_local int g_status = 0;
void set_status(int rc) { g_status = rc; }

static std::atomic<int> g_timing{0};
constexpr int TIMER_RING = 2048;
static thread_local cudaEvent_t g_ev[TIMER_RING][2];
static thread_local int g_ev_created = 0, g_ev_count = 0;

MainKernelTimer::MainKernelTimer(cudaStream_t s) : st(s), on(g_timing.load() != 0) {
    if (!on) return;
    if (g_ev_count >= TIMER_RING) { on = false; return; }
    if (g_ev_count >= g_ev_created) {
        cudaEventCreate(&g_ev[g_ev_created][0]);
        cudaEventCreate(&g_ev[g_ev_created][1]);
        g_ev_created++;
    }
    cudaEventRecord(g_ev[g_ev_count][0], st);
}
MainKernelTimer::~MainKernelTimer() {
    if (!on) return;
    cudaEventRecord(g_ev[g_ev_count][1], st);
    g_ev_count++;
}

static std::atomic<int> g_reserved_sms{0};
int reserved_sms() {
    // BPMF_RESERVED_SMS overrides bpmf_set_reserved_sms() (tuning experiments: fewer SMs under the power cap)
    if (const char *e = getenv("BPMF_RESERVED_SMS")) return atoi(e) < 0 ? 0 : atoi(e);
    return g_reserved_sms.load();
}

int num_sms() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

static MfShape make_shape(size_t L, size_t N, size_t T, size_t S, size_t C, size_t step,
                          int normalize, int network_sum) {
    MfShape sh;
    sh.T = T; sh.S = S; sh.C = C; sh.L = L; sh.N = N; sh.step = step;
    sh.n_corr = bpmf_mf_n_corr(N, L, step);
    sh.normalize = normalize; sh.network_sum = network_sum;
    return sh;
}

// fp32 path staging: window energy (+ mean) per channel on the trace side, template statistics and
// bounds on the template side
struct MfTraceWs {
    float *energy, *mean;
};
static MfTraceWs carve_fp32_traces(Workspace &ws, const MfShape &sh) {
    MfTraceWs w;
    const size_t n_win = sh.N >= sh.L ? sh.N - sh.L + 1 : 0;
    w.energy = ws.take<float>(sh.SC() * n_win);
    w.mean = sh.normalize ? ws.take<float>(sh.SC() * n_win) : nullptr;
    return w;
}
static MfMeta carve_meta(Workspace &ws, const MfShape &sh) {
    MfMeta m;
    m.sst = ws.take<float>(sh.T * sh.SC());
    m.sumt = ws.take<float>(sh.T * sh.SC());
    m.lo = ws.take<int>(sh.T);
    m.hi = ws.take<int>(sh.T);
    return m;
}

// Shape-only resolution: for AUTO this is the FASTEST arithmetic the shape allows; the call itself
// may still pick the split variant from the weights (mf_auto_precision).
static int resolve_precision(const MfShape &sh, int precision) {
    const bool tc_ok = mf_tc_supported(sh.L, sh.step, sh.normalize, sh.network_sum);
    if (precision == BPMF_MF_AUTO) return tc_ok ? BPMF_MF_TC_FP16 : BPMF_MF_FP32;
    if ((precision == BPMF_MF_TC_FP16 || precision == BPMF_MF_TC_FP16X3) && !tc_ok) return -1;
    if (precision < BPMF_MF_AUTO || precision > BPMF_MF_TC_FP16X3) return -1;
    return precision;
}

// AUTO on a tensor-capable shape, decided PER TEMPLATE: single-pass fp16 operands carry 2^-11 relative
// rounding; a template's CC sum then errs by about 4e-4 * sqrt(sum_ch w^2) / sqrt(L) rms.  A template takes the
// single pass when 6.5 sigma of that stays below 6e-5 (tolerance 1e-4 with margin), i.e. when
// sum_ch w^2 < (6e-5 / 2.6e-3)^2 * L; otherwise the hi/lo split (fp32 grade).  One template with concentrated
// weights therefore no longer drags a whole call of well-spread templates into the 3x slower arithmetic.
__host__ __device__ inline float mf_auto_w2_limit(size_t L) {
    const double r = 6e-5 / 2.6e-3;
    return (float)(r * r * (double)L);
}

// one warp per template: sum_ch w^2 -> class; w_easy / w_hard are the weights with the other class zeroed
__global__ void mf_weight_class_kernel(const float *__restrict__ w, size_t T, size_t SC, float w2_limit,
                                       float *__restrict__ w_easy, float *__restrict__ w_hard,
                                       unsigned *__restrict__ counts) {
    const size_t t = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (t >= T) return;
    float a = 0.f;
    for (size_t c = threadIdx.x & 31; c < SC; c += 32) {
        const float v = w[t * SC + c];
        a = fmaf(v, v, a);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    const bool easy = a < w2_limit;
    for (size_t c = threadIdx.x & 31; c < SC; c += 32) {
        const float v = w[t * SC + c];
        w_easy[t * SC + c] = easy ? v : 0.f;
        w_hard[t * SC + c] = easy ? 0.f : v;
    }
    if ((threadIdx.x & 31) == 0 && a > 0.f) atomicAdd(counts + (easy ? 0 : 1), 1u);
}

static thread_local int g_last_precision = 0;
static bool is_tc(int prec) {
    return prec == BPMF_MF_TC_FP16 || prec == BPMF_MF_TC_FP16X3 || prec == BPMF_MF_TC_MIXED;
}

}  // namespace bpmf

using namespace bpmf;

extern "C" const char *bpmf_last_error(void) { return g_err; }
extern "C" int bpmf_last_status(void) { return g_status; }
extern "C" const char *bpmf_version(void) { return "bpmf_b200 r1 sm_100a"; }
extern "C" unsigned long long bpmf_launch_count(void) { return g_launches.load(); }

extern "C" void bpmf_set_kernel_timing(int enable) {
    g_timing.store(enable ? 1 : 0);
    g_ev_count = 0;
}
extern "C" int bpmf_main_kernel_times(float *ms, int capacity) {
    int n = 0;
    for (int i = 0; i < g_ev_count && n < capacity; i++) {
        float t = -1.f;
        if (cudaEventSynchronize(g_ev[i][1]) != cudaSuccess ||
            cudaEventElapsedTime(&t, g_ev[i][0], g_ev[i][1]) != cudaSuccess) {
            cudaGetLastError();
            continue;
        }
        ms[n++] = t;
    }
    g_ev_count = 0;
    return n;
}

extern "C" void bpmf_set_reserved_sms(int n) { g_reserved_sms.store(n < 0 ? 0 : n); }

extern "C" int bpmf_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

// BPMF/GPU.cu:10-24: print the devices visible to the process.
extern "C" void caract_GPU(void) {
    const int n = bpmf_device_count();
    printf("%d cuda devices found on the node\n", n);
    for (int i = 0; i < n; i++) {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, i) != cudaSuccess) continue;
        printf("Device nb %d, name %s (sm_%d%d, %d SMs)\n", i, prop.name, prop.major, prop.minor,
               prop.multiProcessorCount);
        printf("Total memory: %zu \n", (size_t)prop.totalGlobalMem);
        printf("Shared memory per block: %zu (opt-in %zu)\n", (size_t)prop.sharedMemPerBlock,
               (size_t)prop.sharedMemPerBlockOptin);
    }
    fflush(stdout);
}

extern "C" size_t bpmf_mf_n_corr(size_t N, size_t L, size_t step) {
    if (step == 0 || L == 0 || N < L) return 0;
    return (N - L) / step + 1;
}

extern "C" int bpmf_mf_last_precision(void) { return g_last_precision; }

extern "C" int bpmf_mf_resolve_precision(size_t L, size_t step, int normalize, int network_sum,
                                         int precision) {
    MfShape sh = make_shape(L, L, 1, 1, 1, step, normalize, network_sum);
    return resolve_precision(sh, precision);
}

// the split arithmetic needs the residual (lo) traces; AUTO may pick it at run time
static bool wants_lo(int precision) { return precision == BPMF_MF_TC_FP16X3 || precision == BPMF_MF_AUTO; }

extern "C" size_t bpmf_mf_trace_workspace_bytes(size_t L, size_t N, size_t S, size_t C, size_t step,
                                                int normalize, int network_sum, int precision) {
    MfShape sh = make_shape(L, N, 1, S, C, step, normalize, network_sum);
    const int prec = resolve_precision(sh, precision);
    if (is_tc(prec)) return mf_tc_trace_ws_bytes(sh, wants_lo(precision));
    Workspace ws(nullptr, 0);
    carve_fp32_traces(ws, sh);
    return ws.used;
}

extern "C" size_t bpmf_mf_template_workspace_bytes(size_t L, size_t N, size_t T, size_t S, size_t C,
                                                   size_t step, int normalize, int network_sum,
                                                   int precision) {
    MfShape sh = make_shape(L, N, T, S, C, step, normalize, network_sum);
    const int prec = resolve_precision(sh, precision);
    Workspace ws(nullptr, 0);
    ws.take<float>(64);  // scratch of the AUTO policy
    carve_meta(ws, sh);
    if (is_tc(prec)) {
        if (precision == BPMF_MF_AUTO) {  // per-template class: the two masked copies of the weights
            ws.take<float>(sh.T * sh.SC());
            ws.take<float>(sh.T * sh.SC());
        }
        size_t need = mf_tc_template_ws_bytes(sh, prec == BPMF_MF_TC_FP16X3);
        if (precision == BPMF_MF_AUTO) need = std::max(need, mf_tc_template_ws_bytes(sh, 1));
        return ws.used + need;
    }
    return ws.used;
}

extern "C" size_t bpmf_mf_workspace_bytes(size_t L, size_t N, size_t T, size_t S, size_t C,
                                          size_t step, int normalize, int network_sum,
                                          int precision) {
    return align_up(bpmf_mf_trace_workspace_bytes(L, N, S, C, step, normalize, network_sum, precision), 256) +
           bpmf_mf_template_workspace_bytes(L, N, T, S, C, step, normalize, network_sum, precision);
}

extern "C" int bpmf_mf_prepare_traces_dev(const float *d_data, size_t step, size_t L, size_t N, size_t S,
                                          size_t C, int normalize, int network_sum, int precision,
                                          void *d_trace_ws, size_t trace_ws_bytes, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    BPMF_REQUIRE(step >= 1 && L >= 1 && S >= 1 && C >= 1, "matched_filter: empty shape");
    BPMF_REQUIRE(N >= L, "matched_filter: data (%zu) shorter than template (%zu)", N, L);
    BPMF_REQUIRE(N < ((size_t)1 << 31), "matched_filter: n_samples_data exceeds int32 indexing");
    MfShape sh = make_shape(L, N, 1, S, C, step, normalize ? 1 : 0, network_sum ? 1 : 0);
    const int prec = resolve_precision(sh, precision);
    BPMF_REQUIRE(prec > 0, "matched_filter: precision %d is not available for L=%zu step=%zu "
                 "normalize=%d network_sum=%d", precision, L, step, normalize, network_sum);
    Workspace ws(d_trace_ws, trace_ws_bytes);
    if (is_tc(prec)) return mf_tc_prepare_traces(sh, wants_lo(precision), d_data, ws, st);
    MfTraceWs w = carve_fp32_traces(ws, sh);
    BPMF_REQUIRE(ws.ok(), "matched_filter: trace workspace too small (%zu < %zu)", trace_ws_bytes, ws.used);
    return mf_energy(sh, d_data, w.energy, w.mean, sh.N - sh.L + 1, false, false, st);
}

extern "C" int bpmf_mf_run_dev(const float *d_templates, const int *d_moveouts, const float *d_weights,
                               const float *d_data, size_t step, size_t L, size_t N, size_t T, size_t S,
                               size_t C, int normalize, int network_sum, int precision, void *d_trace_ws,
                               size_t trace_ws_bytes, void *d_template_ws, size_t template_ws_bytes,
                               float *d_cc, size_t cc_pitch, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    BPMF_REQUIRE(step >= 1 && L >= 1 && S >= 1 && C >= 1, "matched_filter: empty shape");
    BPMF_REQUIRE(N >= L, "matched_filter: data (%zu) shorter than template (%zu)", N, L);
    if (T == 0) return BPMF_OK;
    MfShape sh = make_shape(L, N, T, S, C, step, normalize ? 1 : 0, network_sum ? 1 : 0);
    BPMF_REQUIRE(!sh.network_sum || cc_pitch >= sh.n_corr, "matched_filter: cc_pitch < n_corr");
    int prec = resolve_precision(sh, precision);
    BPMF_REQUIRE(prec > 0, "matched_filter: precision %d is not available for L=%zu step=%zu "
                 "normalize=%d network_sum=%d", precision, L, step, normalize, network_sum);
    BPMF_REQUIRE(d_trace_ws && d_template_ws, "matched_filter: workspace missing");
    // the tensor path accumulates with 8/16-byte vector reductions: rows of d_cc must start 16-byte aligned
    BPMF_REQUIRE(!is_tc(prec) || ((cc_pitch & 3) == 0 && ((uintptr_t)d_cc & 15) == 0),
                 "matched_filter: the tensor path needs d_cc 16-byte aligned and cc_pitch a multiple of 4 "
                 "(got %p, %zu); pad the pitch (e.g. to 64 elements) or ask for BPMF_MF_FP32", (void *)d_cc, cc_pitch);
    Workspace tws(d_trace_ws, trace_ws_bytes), pws(d_template_ws, template_ws_bytes);
    float *scratch = pws.take<float>(64);
    MfMeta meta = carve_meta(pws, sh);
    BPMF_REQUIRE(pws.ok(), "matched_filter: template workspace too small");
    int rc;
    if (precision == BPMF_MF_AUTO && prec == BPMF_MF_TC_FP16) {
        // classify every template on the device; the only host synchronisation of this entry point reads back
        // the two class counts (8 bytes)
        float *w_easy = pws.take<float>(sh.T * sh.SC());
        float *w_hard = pws.take<float>(sh.T * sh.SC());
        BPMF_REQUIRE(pws.ok(), "matched_filter: template workspace too small");
        unsigned *counts = reinterpret_cast<unsigned *>(scratch);
        BPMF_CUDA(cudaMemsetAsync(counts, 0, 2 * sizeof(unsigned), st));
        mf_weight_class_kernel<<<(unsigned)((sh.T + 7) / 8), 256, 0, st>>>(d_weights, sh.T, sh.SC(),
                                                                         mf_auto_w2_limit(sh.L), w_easy, w_hard, counts);
        BPMF_LAUNCH_CHECK();
        unsigned h_counts[2] = {0, 0};
        BPMF_CUDA(cudaMemcpyAsync(h_counts, counts, sizeof(h_counts), cudaMemcpyDeviceToHost, st));
        BPMF_CUDA(cudaStreamSynchronize(st));
        if (h_counts[1] == 0) prec = BPMF_MF_TC_FP16;
        else if (h_counts[0] == 0) prec = BPMF_MF_TC_FP16X3;
        else {
            // both classes present: the single pass over the well-spread templates, then the split over the
            // others, accumulating into the same slab (each template row is touched by exactly one of them;
            // tiles without a weighted template of the class are skipped by the kernel)
            g_last_precision = BPMF_MF_TC_MIXED;
            if ((rc = mf_prepass(sh, d_templates, d_moveouts, w_easy, d_data, meta, nullptr, nullptr, st)) != BPMF_OK)
                return rc;
            Workspace p1 = pws;
            if ((rc = mf_tc_run(sh, 0, d_templates, d_moveouts, w_easy, meta, tws, true, p1, d_cc, cc_pitch, true,
                                st)) != BPMF_OK)
                return rc;
            if ((rc = mf_prepass(sh, d_templates, d_moveouts, w_hard, d_data, meta, nullptr, nullptr, st)) != BPMF_OK)
                return rc;
            Workspace p2 = pws, t2(d_trace_ws, trace_ws_bytes);
            return mf_tc_run(sh, 1, d_templates, d_moveouts, w_hard, meta, t2, true, p2, d_cc, cc_pitch, false, st);
        }
    }
    g_last_precision = prec;
    if ((rc = mf_prepass(sh, d_templates, d_moveouts, d_weights, d_data, meta, nullptr, nullptr, st)) != BPMF_OK)
        return rc;
    if (is_tc(prec))
        return mf_tc_run(sh, prec == BPMF_MF_TC_FP16X3, d_templates, d_moveouts, d_weights, meta, tws,
                         wants_lo(precision), pws, d_cc, cc_pitch, true, st);
    MfTraceWs w = carve_fp32_traces(tws, sh);
    BPMF_REQUIRE(tws.ok(), "matched_filter: trace workspace too small");
    if (!sh.network_sum) BPMF_CUDA(cudaMemsetAsync(d_cc, 0, sizeof(float) * sh.T * sh.n_corr * sh.SC(), st));
    return mf_fp32_run(sh, d_templates, d_moveouts, d_weights, d_data, meta, w.energy, w.mean, d_cc, cc_pitch, st);
}

extern "C" int bpmf_matched_filter_dev(const float *d_templates, const int *d_moveouts,
                                       const float *d_weights, const float *d_data, size_t step,
                                       size_t L, size_t N, size_t T, size_t S, size_t C,
                                       int normalize, int network_sum, int precision,
                                       void *d_workspace, size_t workspace_bytes, float *d_cc,
                                       size_t cc_pitch, void *stream) {
    if (T == 0) return BPMF_OK;
    const size_t tb = align_up(bpmf_mf_trace_workspace_bytes(L, N, S, C, step, normalize, network_sum, precision), 256);
    const size_t pb = bpmf_mf_template_workspace_bytes(L, N, T, S, C, step, normalize, network_sum, precision);
    BPMF_REQUIRE(d_workspace && workspace_bytes >= tb + pb, "matched_filter: workspace too small (%zu < %zu)",
                 workspace_bytes, tb + pb);
    int rc = bpmf_mf_prepare_traces_dev(d_data, step, L, N, S, C, normalize, network_sum, precision, d_workspace,
                                        tb, stream);
    if (rc != BPMF_OK) return rc;
    return bpmf_mf_run_dev(d_templates, d_moveouts, d_weights, d_data, step, L, N, T, S, C, normalize, network_sum,
                           precision, d_workspace, tb, (char *)d_workspace + tb, pb, d_cc, cc_pitch, stream);
}

// ---------------------------------------------------------------- host-pointer entry point
// Device buffers of the host path are cached per device (grow-only) so that the repeated calls BPMF makes
// -- one per template chunk and per day of traces, BPMF/similarity_search.py:783-790 -- do not pay
// cudaMalloc/cudaFree every time.  bpmf_release_cache() frees them.
namespace bpmf {
struct HostCache {
    void *ptr[16] = {nullptr};
    size_t cap[16] = {0};
    cudaStream_t st[2] = {nullptr, nullptr};
    cudaEvent_t ev = nullptr;
    bool busy = false;
    // bpmf_mf_hold_traces: the caller promises that this host buffer does not change until it is released, so the
    // uploaded + staged copy (slots 0 and 1) can serve every call made on it meanwhile
    const float *held_ptr = nullptr;
    size_t held_N = 0, held_SC = 0;
    bool staged = false;
    size_t st_L = 0, st_step = 0;
    int st_norm = 0, st_netsum = 0, st_prec = 0;
};
static HostCache g_cache[16];
static std::mutex g_cache_mu;

static cudaError_t cache_get(HostCache &c, int slot, size_t bytes, void **out) {
    if (c.cap[slot] < bytes) {
        if (slot <= 1) c.staged = false;  // the traces (slot 0) or their staging (slot 1) move
        if (c.ptr[slot]) cudaFree(c.ptr[slot]);
        c.ptr[slot] = nullptr;
        c.cap[slot] = 0;
        cudaError_t e = cudaMalloc(&c.ptr[slot], bytes ? bytes : 1);
        if (e != cudaSuccess) return e;
        c.cap[slot] = bytes;
    }
    *out = c.ptr[slot];
    return cudaSuccess;
}
static void cache_release(HostCache &c) {
    for (int i = 0; i < 16; i++) {
        if (c.ptr[i]) cudaFree(c.ptr[i]);
        c.ptr[i] = nullptr;
        c.cap[i] = 0;
    }
    c.staged = false;  // a held day has to be uploaded and staged again
}
}  // namespace bpmf

extern "C" void bpmf_release_cache(void) {
    host_pool().release();  // the calling thread's libc scratch
    std::lock_guard<std::mutex> lk(g_cache_mu);
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return;
    if (!g_cache[dev].busy) cache_release(g_cache[dev]);
}

// BPMF runs its template groups chunk by chunk over ONE day of traces (BPMF/similarity_search.py:773-790: every chunk
// is a separate fast_matched_filter.matched_filter call on the same `data` array).  Between bpmf_mf_hold_traces and
// bpmf_mf_release_traces the host entry point uploads and stages that array once and reuses it for every call that
// passes the same pointer and shape, instead of moving 1.7-3.3 GB per chunk.  Explicit on purpose: a pointer can be
// reused for the next day's samples, so nothing is cached unless the caller says the buffer is stable.
extern "C" int bpmf_mf_hold_traces(const float *data, size_t N, size_t S, size_t C) {
    int dev = 0;
    BPMF_CUDA(cudaGetDevice(&dev));
    BPMF_REQUIRE(dev >= 0 && dev < 16, "hold_traces: device index out of range");
    std::lock_guard<std::mutex> lk(g_cache_mu);
    HostCache &c = g_cache[dev];
    BPMF_REQUIRE(!c.busy, "hold_traces: a matched-filter call is running on this device");
    c.held_ptr = data;
    c.held_N = N;
    c.held_SC = S * C;
    c.staged = false;
    return BPMF_OK;
}

extern "C" void bpmf_mf_release_traces(void) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return;
    std::lock_guard<std::mutex> lk(g_cache_mu);
    g_cache[dev].held_ptr = nullptr;
    g_cache[dev].staged = false;
}

// Traces go up and are staged once; templates are processed in chunks on two streams so the
// device->host copy of one chunk's CC slab overlaps the next chunk's kernels (the caller's buffers
// may be pinned -- then the copies are truly asynchronous -- or pageable).
extern "C" int bpmf_matched_filter(const float *templates, const int *moveouts, const float *weights,
                                   const float *data, size_t step, size_t L, size_t N, size_t T,
                                   size_t S, size_t C, int normalize, int network_sum, int precision,
                                   float *cc_out) {
    BPMF_REQUIRE(step >= 1 && L >= 1 && S >= 1 && C >= 1, "matched_filter: empty shape");
    BPMF_REQUIRE(N >= L, "matched_filter: data (%zu) shorter than template (%zu)", N, L);
    if (T == 0) return BPMF_OK;
    const size_t SC = S * C;
    const size_t n_corr = bpmf_mf_n_corr(N, L, step);
    const size_t pitch = network_sum ? align_up(n_corr, 64) : 0;
    const size_t row_elems = network_sum ? pitch : n_corr * SC;

    // AUTO is resolved per template (same rule as mf_weight_class_kernel), on the host copy of the weights: when every
    // template falls in one class the chunk loop below never synchronises; a mixed set keeps AUTO and lets the
    // device entry point run the two classes separately (8 bytes read back per chunk)
    if (precision == BPMF_MF_AUTO) {
        MfShape sh0 = make_shape(L, N, 1, S, C, step, normalize ? 1 : 0, network_sum ? 1 : 0);
        if (resolve_precision(sh0, precision) == BPMF_MF_TC_FP16) {
            const double lim = (double)mf_auto_w2_limit(L);
            size_t n_easy = 0, n_hard = 0;
            for (size_t t = 0; t < T; t++) {
                float a = 0.f;
                for (size_t q = 0; q < SC; q++) a = fmaf(weights[t * SC + q], weights[t * SC + q], a);
                if (a > 0.f) ((double)a < lim ? n_easy : n_hard)++;
            }
            if (n_hard == 0) precision = BPMF_MF_TC_FP16;
            else if (n_easy == 0) precision = BPMF_MF_TC_FP16X3;
        }
    }

    int dev = 0;
    BPMF_CUDA(cudaGetDevice(&dev));
    // one host-path call at a time per device uses the cache; a concurrent caller gets private buffers
    HostCache local;
    HostCache *c = &local;
    bool cached = false;
    if (dev >= 0 && dev < 16) {
        std::lock_guard<std::mutex> lk(g_cache_mu);
        if (!g_cache[dev].busy) {
            g_cache[dev].busy = true;
            c = &g_cache[dev];
            cached = true;
        }
    }
    int rc = BPMF_OK;
    auto finish = [&](int code) {
        if (cached) {
            std::lock_guard<std::mutex> lk(g_cache_mu);
            c->busy = false;
        } else {
            cache_release(local);
            for (int i = 0; i < 2; i++)
                if (local.st[i]) cudaStreamDestroy(local.st[i]);
            if (local.ev) cudaEventDestroy(local.ev);
        }
        return code;
    };
#define MF_TRY(expr)                                                                        \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess) {                                                            \
            set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            cudaDeviceSynchronize();                                                        \
            cudaGetLastError();                                                             \
            return finish(BPMF_E_CUDA);                                                     \
        }                                                                                   \
    } while (0)
    for (int i = 0; i < 2; i++)
        if (!c->st[i]) MF_TRY(cudaStreamCreateWithFlags(&c->st[i], cudaStreamNonBlocking));
    if (!c->ev) MF_TRY(cudaEventCreateWithFlags(&c->ev, cudaEventDisableTiming));

    const size_t trace_b = bpmf_mf_trace_workspace_bytes(L, N, S, C, step, normalize, network_sum, precision);
    void *d_data = nullptr, *d_trace = nullptr;
    MF_TRY(cache_get(*c, 0, sizeof(float) * SC * N, &d_data));
    MF_TRY(cache_get(*c, 1, trace_b, &d_trace));
    size_t free_b = 0, total_b = 0;
    MF_TRY(cudaMemGetInfo(&free_b, &total_b));
    for (int i = 2; i < 12; i++) free_b += c->cap[i];  // chunk buffers already held by the cache are reusable

    // chunk size: ~16 chunks when T is large (the first chunk's kernels are the only ones the device->host
    // copies cannot hide behind), bounded by half of the free memory
    size_t Tc = T;
    if (T >= 64) Tc = std::max<size_t>(32, align_up((T + 15) / 16, 16));
    auto need = [&](size_t tc) {
        return 2 * (sizeof(float) * tc * row_elems + sizeof(float) * tc * SC * L + 8 * tc * SC +
                    bpmf_mf_template_workspace_bytes(L, N, tc, S, C, step, normalize, network_sum, precision));
    };
    while (Tc > 16 && need(Tc) > free_b / 2) Tc = align_up(Tc / 2, 16);
    while (Tc > 1 && need(Tc) > free_b / 2) Tc = (Tc + 1) / 2;
    if (need(Tc) > free_b) {
        set_error("matched_filter: not enough device memory (%zu needed, %zu free)", need(Tc), free_b);
        return finish(BPMF_E_NOMEM);
    }
    const size_t tpl_b = bpmf_mf_template_workspace_bytes(L, N, Tc, S, C, step, normalize, network_sum, precision);
    const int n_buf = (Tc < T) ? 2 : 1;
    void *d_tp[2], *d_mv[2], *d_w[2], *d_cc[2], *d_ws[2];
    for (int i = 0; i < n_buf; i++) {
        MF_TRY(cache_get(*c, 2 + 5 * i + 0, sizeof(float) * Tc * SC * L, &d_tp[i]));
        MF_TRY(cache_get(*c, 2 + 5 * i + 1, sizeof(int) * Tc * SC, &d_mv[i]));
        MF_TRY(cache_get(*c, 2 + 5 * i + 2, sizeof(float) * Tc * SC, &d_w[i]));
        MF_TRY(cache_get(*c, 2 + 5 * i + 3, sizeof(float) * Tc * row_elems, &d_cc[i]));
        MF_TRY(cache_get(*c, 2 + 5 * i + 4, tpl_b, &d_ws[i]));
    }
    const bool held = cached && c->held_ptr == data && c->held_N == N && c->held_SC == SC;
    const bool reuse = held && c->staged && c->st_L == L && c->st_step == step && c->st_norm == normalize &&
                       c->st_netsum == network_sum && c->st_prec == precision;
    if (!reuse) {
        c->staged = false;
        MF_TRY(cudaMemcpyAsync(d_data, data, sizeof(float) * SC * N, cudaMemcpyHostToDevice, c->st[0]));
        rc = bpmf_mf_prepare_traces_dev((const float *)d_data, step, L, N, S, C, normalize, network_sum, precision,
                                        d_trace, trace_b, c->st[0]);
        if (rc != BPMF_OK) { cudaDeviceSynchronize(); return finish(rc); }
        if (held) {
            c->staged = true;
            c->st_L = L; c->st_step = step; c->st_norm = normalize; c->st_netsum = network_sum; c->st_prec = precision;
        }
    }
    MF_TRY(cudaEventRecord(c->ev, c->st[0]));

    size_t chunk = 0;
    for (size_t t0 = 0; t0 < T; t0 += Tc, chunk++) {
        const size_t tc = std::min(Tc, T - t0);
        const int b = (int)(chunk % n_buf);
        cudaStream_t s = c->st[b];
        if (chunk == 1) MF_TRY(cudaStreamWaitEvent(s, c->ev, 0));
        MF_TRY(cudaMemcpyAsync(d_tp[b], templates + t0 * SC * L, sizeof(float) * tc * SC * L,
                               cudaMemcpyHostToDevice, s));
        MF_TRY(cudaMemcpyAsync(d_mv[b], moveouts + t0 * SC, sizeof(int) * tc * SC, cudaMemcpyHostToDevice, s));
        MF_TRY(cudaMemcpyAsync(d_w[b], weights + t0 * SC, sizeof(float) * tc * SC, cudaMemcpyHostToDevice, s));
        rc = bpmf_mf_run_dev((const float *)d_tp[b], (const int *)d_mv[b], (const float *)d_w[b],
                             (const float *)d_data, step, L, N, tc, S, C, normalize, network_sum, precision,
                             d_trace, trace_b, d_ws[b], tpl_b, (float *)d_cc[b], pitch, s);
        if (rc != BPMF_OK) break;
        if (network_sum)
            MF_TRY(cudaMemcpy2DAsync(cc_out + t0 * n_corr, sizeof(float) * n_corr, d_cc[b],
                                     sizeof(float) * pitch, sizeof(float) * n_corr, tc,
                                     cudaMemcpyDeviceToHost, s));
        else
            MF_TRY(cudaMemcpyAsync(cc_out + t0 * row_elems, d_cc[b], sizeof(float) * tc * row_elems,
                                   cudaMemcpyDeviceToHost, s));
    }
    for (int i = 0; i < 2; i++) {
        cudaError_t e = cudaStreamSynchronize(c->st[i]);
        if (e != cudaSuccess && rc == BPMF_OK) {
            set_error("matched_filter: stream sync -> %s", cudaGetErrorString(e));
            cudaGetLastError();
            rc = BPMF_E_CUDA;
        }
    }
#undef MF_TRY
    return finish(rc);
}

// Upstream fast_matched_filter symbols [UPSTREAM-UNVERIFIED]: void return, errors reported on stderr.
extern "C" void matched_filter(float *templates, float *sum_square_templates, int *moveouts,
                               float *data, float *weights, size_t step, size_t L, size_t N, size_t T,
                               size_t S, size_t C, size_t n_corr, int normalize, float *cc_sum) {
    (void)sum_square_templates;
    (void)n_corr;
    const int rc = bpmf_matched_filter(templates, moveouts, weights, data, step, L, N, T, S, C, normalize, 1,
                                       BPMF_MF_AUTO, cc_sum);
    set_status(rc);
    if (rc != BPMF_OK) fprintf(stderr, "libbpmf_b200: matched_filter failed: %s\n", bpmf_last_error());
}

extern "C" void matched_filter_no_sum(float *templates, float *sum_square_templates, int *moveouts,
                                      float *data, float *weights, size_t step, size_t L, size_t N,
                                      size_t T, size_t S, size_t C, size_t n_corr, int normalize,
                                      float *cc) {
    (void)sum_square_templates;
    (void)n_corr;
    const int rc = bpmf_matched_filter(templates, moveouts, weights, data, step, L, N, T, S, C, normalize, 0,
                                       BPMF_MF_AUTO, cc);
    set_status(rc);
    if (rc != BPMF_OK) fprintf(stderr, "libbpmf_b200: matched_filter_no_sum failed: %s\n", bpmf_last_error());
}
