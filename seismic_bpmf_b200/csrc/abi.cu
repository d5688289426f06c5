// abi.cu -- extern "C" surface of libbpmf_b200.so: error plumbing, device query and the
// matched-filter entry points (include/bpmf_b200.h groups (2a) and (3)).
#include <stdarg.h>
#include <string.h>

#include <algorithm>
#include <vector>

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

static std::atomic<int> g_timing{0};
constexpr int TIMER_RING = 256;
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

struct MfWs {
    MfMeta meta;
    float *energy, *mean;
};

static MfWs carve_fp32(Workspace &ws, const MfShape &sh) {
    MfWs w;
    w.meta.sst = ws.take<float>(sh.T * sh.SC());
    w.meta.sumt = ws.take<float>(sh.T * sh.SC());
    w.meta.lo = ws.take<int>(sh.T);
    w.meta.hi = ws.take<int>(sh.T);
    const size_t n_win = sh.N >= sh.L ? sh.N - sh.L + 1 : 0;
    w.energy = ws.take<float>(sh.SC() * n_win);
    w.mean = sh.normalize ? ws.take<float>(sh.SC() * n_win) : nullptr;
    return w;
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

// max over templates of sum_ch w^2 (one warp per template)
__global__ void mf_weight_norm_kernel(const float *__restrict__ w, size_t T, size_t SC, float *__restrict__ out) {
    const size_t t = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (t >= T) return;
    float a = 0.f;
    for (size_t c = threadIdx.x & 31; c < SC; c += 32) {
        const float v = w[t * SC + c];
        a = fmaf(v, v, a);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if ((threadIdx.x & 31) == 0 && a > 0.f) atomicMax(reinterpret_cast<unsigned *>(out), __float_as_uint(a));
}

static thread_local int g_last_precision = 0;

// AUTO on a tensor-capable shape: single-pass fp16 operands carry 2^-11 relative rounding; the CC
// sum then errs by about 4e-4 * sqrt(sum_ch w^2) / sqrt(L) rms.  Take the single pass when 6.5
// sigma of that stays below 6e-5 (tolerance 1e-4 with margin), else the hi/lo split (fp32 grade).
static int mf_auto_precision(const MfShape &sh, const float *d_weights, float *d_scratch, cudaStream_t st,
                             int *out) {
    BPMF_CUDA(cudaMemsetAsync(d_scratch, 0, sizeof(float), st));
    mf_weight_norm_kernel<<<(unsigned)((sh.T + 7) / 8), 256, 0, st>>>(d_weights, sh.T, sh.SC(), d_scratch);
    BPMF_LAUNCH_CHECK();
    float w2 = 0.f;
    BPMF_CUDA(cudaMemcpyAsync(&w2, d_scratch, sizeof(float), cudaMemcpyDeviceToHost, st));
    BPMF_CUDA(cudaStreamSynchronize(st));
    const double predicted = 2.6e-3 * sqrt((double)w2) / sqrt((double)sh.L);
    *out = (predicted < 6e-5) ? BPMF_MF_TC_FP16 : BPMF_MF_TC_FP16X3;
    return BPMF_OK;
}

}  // namespace bpmf

using namespace bpmf;

extern "C" const char *bpmf_last_error(void) { return g_err; }
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

extern "C" size_t bpmf_mf_workspace_bytes(size_t L, size_t N, size_t T, size_t S, size_t C,
                                          size_t step, int normalize, int network_sum,
                                          int precision) {
    MfShape sh = make_shape(L, N, T, S, C, step, normalize, network_sum);
    const int prec = resolve_precision(sh, precision);
    Workspace ws(nullptr, 0);
    if (prec == BPMF_MF_TC_FP16 || prec == BPMF_MF_TC_FP16X3) {
        ws.take<float>(sh.T * sh.SC());
        ws.take<float>(sh.T * sh.SC());
        ws.take<int>(sh.T);
        ws.take<int>(sh.T);
        size_t need = mf_tc_workspace_bytes(sh, prec == BPMF_MF_TC_FP16X3);
        if (precision == BPMF_MF_AUTO) need = std::max(need, mf_tc_workspace_bytes(sh, 1));  // AUTO may split
        return ws.used + need;
    }
    carve_fp32(ws, sh);
    return ws.used;
}

extern "C" int bpmf_matched_filter_dev(const float *d_templates, const int *d_moveouts,
                                       const float *d_weights, const float *d_data, size_t step,
                                       size_t L, size_t N, size_t T, size_t S, size_t C,
                                       int normalize, int network_sum, int precision,
                                       void *d_workspace, size_t workspace_bytes, float *d_cc,
                                       size_t cc_pitch, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    BPMF_REQUIRE(step >= 1 && L >= 1 && S >= 1 && C >= 1, "matched_filter: empty shape");
    BPMF_REQUIRE(N >= L, "matched_filter: data (%zu) shorter than template (%zu)", N, L);
    BPMF_REQUIRE(N < ((size_t)1 << 31), "matched_filter: n_samples_data exceeds int32 indexing");
    if (T == 0) return BPMF_OK;
    MfShape sh = make_shape(L, N, T, S, C, step, normalize ? 1 : 0, network_sum ? 1 : 0);
    BPMF_REQUIRE(!sh.network_sum || cc_pitch >= sh.n_corr, "matched_filter: cc_pitch < n_corr");
    int prec = resolve_precision(sh, precision);
    BPMF_REQUIRE(prec > 0, "matched_filter: precision %d is not available for L=%zu step=%zu "
                 "normalize=%d network_sum=%d", precision, L, step, normalize, network_sum);
    BPMF_REQUIRE(workspace_bytes >= 256 && d_workspace != nullptr, "matched_filter: workspace missing");
    Workspace ws(d_workspace, workspace_bytes);
    int rc;
    if (precision == BPMF_MF_AUTO && prec == BPMF_MF_TC_FP16) {
        // the only host synchronisation of this entry point: 4 bytes of weight statistics
        if ((rc = mf_auto_precision(sh, d_weights, (float *)d_workspace, st, &prec)) != BPMF_OK) return rc;
    }
    g_last_precision = prec;
    if (prec == BPMF_MF_TC_FP16 || prec == BPMF_MF_TC_FP16X3) {
        MfMeta meta;
        meta.sst = ws.take<float>(sh.T * sh.SC());
        meta.sumt = ws.take<float>(sh.T * sh.SC());
        meta.lo = ws.take<int>(sh.T);
        meta.hi = ws.take<int>(sh.T);
        BPMF_REQUIRE(ws.used + mf_tc_workspace_bytes(sh, prec == BPMF_MF_TC_FP16X3) <= workspace_bytes,
                     "matched_filter: workspace too small");
        if ((rc = mf_prepass(sh, d_templates, d_moveouts, d_weights, d_data, meta, nullptr, nullptr,
                             st)) != BPMF_OK)
            return rc;
        return mf_tc_run(sh, prec == BPMF_MF_TC_FP16X3, d_templates, d_moveouts, d_weights, d_data,
                         meta, ws, d_cc, cc_pitch, st);
    }
    MfWs w = carve_fp32(ws, sh);
    BPMF_REQUIRE(ws.ok(), "matched_filter: workspace too small (%zu < %zu)", workspace_bytes, ws.used);
    if ((rc = mf_prepass(sh, d_templates, d_moveouts, d_weights, d_data, w.meta, w.energy, w.mean,
                         st)) != BPMF_OK)
        return rc;
    if (!sh.network_sum) {
        BPMF_CUDA(cudaMemsetAsync(d_cc, 0, sizeof(float) * sh.T * sh.n_corr * sh.SC(), st));
    }
    return mf_fp32_run(sh, d_templates, d_moveouts, d_weights, d_data, w.meta, w.energy, w.mean, d_cc,
                       cc_pitch, st);
}

// Host-pointer entry point.  Data goes up once; templates are processed in chunks on two
// streams so the device->host copy of one chunk's CC slab overlaps the next chunk's kernels
// (the caller's buffers may be pinned -- then the copies are truly asynchronous -- or pageable).
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

    DevBuf d_data;
    BPMF_CUDA(d_data.alloc(sizeof(float) * SC * N));
    size_t free_b = 0, total_b = 0;
    BPMF_CUDA(cudaMemGetInfo(&free_b, &total_b));

    // chunk size: >= 4 chunks for overlap when T is large, bounded by a third of the free memory
    size_t Tc = T;
    if (T >= 64) Tc = (T + 3) / 4;
    auto need = [&](size_t tc) {
        return 2 * (sizeof(float) * tc * row_elems + sizeof(float) * tc * SC * L + 8 * tc * SC) +
               bpmf_mf_workspace_bytes(L, N, tc, S, C, step, normalize, network_sum, precision);
    };
    while (Tc > 1 && need(Tc) > free_b / 3 * 2) Tc = (Tc + 1) / 2;
    BPMF_REQUIRE(need(Tc) <= free_b, "matched_filter: not enough device memory (%zu needed, %zu free)",
                 need(Tc), free_b);
    if (network_sum && Tc >= 32) Tc = Tc / 16 * 16;  // whole tensor-path template tiles per chunk
    const size_t ws_bytes = bpmf_mf_workspace_bytes(L, N, Tc, S, C, step, normalize, network_sum, precision);

    cudaStream_t st[2] = {nullptr, nullptr};
    cudaEvent_t ev_data = nullptr;
    DevBuf d_tp[2], d_mv[2], d_w[2], d_cc[2], d_ws[2];
    int rc = BPMF_OK;
    auto cleanup = [&]() {
        for (int i = 0; i < 2; i++)
            if (st[i]) cudaStreamDestroy(st[i]);
        if (ev_data) cudaEventDestroy(ev_data);
    };
#define MF_TRY(expr)                                                                        \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess) {                                                            \
            set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            cudaDeviceSynchronize();                                                        \
            cleanup();                                                                      \
            return BPMF_E_CUDA;                                                             \
        }                                                                                   \
    } while (0)
    const int n_buf = (Tc < T) ? 2 : 1;
    for (int i = 0; i < n_buf; i++) {
        MF_TRY(cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking));
        MF_TRY(d_tp[i].alloc(sizeof(float) * Tc * SC * L));
        MF_TRY(d_mv[i].alloc(sizeof(int) * Tc * SC));
        MF_TRY(d_w[i].alloc(sizeof(float) * Tc * SC));
        MF_TRY(d_cc[i].alloc(sizeof(float) * Tc * row_elems));
        MF_TRY(d_ws[i].alloc(ws_bytes));
    }
    MF_TRY(cudaEventCreateWithFlags(&ev_data, cudaEventDisableTiming));
    MF_TRY(cudaMemcpyAsync(d_data.p, data, sizeof(float) * SC * N, cudaMemcpyHostToDevice, st[0]));
    MF_TRY(cudaEventRecord(ev_data, st[0]));

    size_t chunk = 0;
    for (size_t t0 = 0; t0 < T; t0 += Tc, chunk++) {
        const size_t tc = std::min(Tc, T - t0);
        const int b = (int)(chunk % n_buf);
        cudaStream_t s = st[b];
        if (chunk == 1) MF_TRY(cudaStreamWaitEvent(s, ev_data, 0));
        MF_TRY(cudaMemcpyAsync(d_tp[b].p, templates + t0 * SC * L, sizeof(float) * tc * SC * L,
                               cudaMemcpyHostToDevice, s));
        MF_TRY(cudaMemcpyAsync(d_mv[b].p, moveouts + t0 * SC, sizeof(int) * tc * SC,
                               cudaMemcpyHostToDevice, s));
        MF_TRY(cudaMemcpyAsync(d_w[b].p, weights + t0 * SC, sizeof(float) * tc * SC,
                               cudaMemcpyHostToDevice, s));
        rc = bpmf_matched_filter_dev(d_tp[b].as<float>(), d_mv[b].as<int>(), d_w[b].as<float>(),
                                     d_data.as<float>(), step, L, N, tc, S, C, normalize, network_sum,
                                     precision, d_ws[b].p, ws_bytes, d_cc[b].as<float>(), pitch, s);
        if (rc != BPMF_OK) break;
        if (network_sum)
            MF_TRY(cudaMemcpy2DAsync(cc_out + t0 * n_corr, sizeof(float) * n_corr, d_cc[b].p,
                                     sizeof(float) * pitch, sizeof(float) * n_corr, tc,
                                     cudaMemcpyDeviceToHost, s));
        else
            MF_TRY(cudaMemcpyAsync(cc_out + t0 * row_elems, d_cc[b].p, sizeof(float) * tc * row_elems,
                                   cudaMemcpyDeviceToHost, s));
    }
    for (int i = 0; i < n_buf; i++) {
        cudaError_t e = cudaStreamSynchronize(st[i]);
        if (e != cudaSuccess && rc == BPMF_OK) {
            set_error("matched_filter: stream sync -> %s", cudaGetErrorString(e));
            rc = BPMF_E_CUDA;
        }
    }
    cleanup();
#undef MF_TRY
    return rc;
}

// Upstream fast_matched_filter symbols [UPSTREAM-UNVERIFIED]: void return, errors reported on stderr.
extern "C" void matched_filter(float *templates, float *sum_square_templates, int *moveouts,
                               float *data, float *weights, size_t step, size_t L, size_t N, size_t T,
                               size_t S, size_t C, size_t n_corr, int normalize, float *cc_sum) {
    (void)sum_square_templates;
    (void)n_corr;
    if (bpmf_matched_filter(templates, moveouts, weights, data, step, L, N, T, S, C, normalize, 1,
                            BPMF_MF_AUTO, cc_sum) != BPMF_OK)
        fprintf(stderr, "libbpmf_b200: matched_filter failed: %s\n", bpmf_last_error());
}

extern "C" void matched_filter_no_sum(float *templates, float *sum_square_templates, int *moveouts,
                                      float *data, float *weights, size_t step, size_t L, size_t N,
                                      size_t T, size_t S, size_t C, size_t n_corr, int normalize,
                                      float *cc) {
    (void)sum_square_templates;
    (void)n_corr;
    if (bpmf_matched_filter(templates, moveouts, weights, data, step, L, N, T, S, C, normalize, 0,
                            BPMF_MF_AUTO, cc) != BPMF_OK)
        fprintf(stderr, "libbpmf_b200: matched_filter_no_sum failed: %s\n", bpmf_last_error());
}
