// intertp.cu -- batched inter-template cross-correlation.
//
// Replaces the loop of BPMF/dataset.py:4784-4830 (TemplateGroup.compute_intertemplate_cc): one
// fast_matched_filter call per template t -- data = the full waveforms of template t (S, C, Lw), templates = the
// waveforms of every other template cut to [max_lag, Lw - max_lag), zero moveouts, network_sum=False -- followed by
// a max over the 2*max_lag + 1 lags per channel and a weighted sum over channels
// (dataset.py:4828-4830).  Here all T x T pairs run in one launch:
//
//   out[t, t2] = sum_{s,c} w[t,s,c] * max_lag CC( X[t2,s,c, max_lag : Lw-max_lag],  X[t,s,c, lag : lag + L] )
//
// with the matched filter's own normalisation (CC = sum T.D / sqrt(sum T^2 * sum D^2), exactly 0
// where sum T^2 * sum D^2 <= 1e-6), zero-weight channels skipped, pairs outside `pair_mask` left at 0.
//
//   intertp_stats_kernel : warp per (template, channel): sum of squares of the cut waveform and the window
//                          energy at every lag (float64 accumulation).
//   intertp_cc_kernel    : CTA = one data template t x 32 templates t2; the channel's data waveform is staged in
//                          shared memory once per CTA, each warp owns 4 of the t2; lanes split the samples in
//                          float4 pieces and slide an 8-lag register window over them (32 FMA per 4 + 3 loads).
//                          Roofline: FP32 FMA pipe, 2 * L * n_lag flop per (pair, channel).
#include "common.cuh"

namespace bpmf {

constexpr int IT_THREADS = 256;
constexpr int IT_WARPS = IT_THREADS / 32;
constexpr int IT_PER_WARP = 4;
constexpr int IT_TB2 = IT_WARPS * IT_PER_WARP;  // templates t2 per CTA
constexpr int IT_LB = 8;                        // lags per register block

__global__ void intertp_stats_kernel(const float *__restrict__ wave, size_t rows, int Lw, int max_lag,
                                     float *__restrict__ sst, float *__restrict__ energy) {
    const size_t row = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const int L = Lw - 2 * max_lag, n_lag = 2 * max_lag + 1;
    const float *x = wave + row * (size_t)Lw;
    for (int lag = 0; lag < n_lag; lag++) {
        double a = 0.0;
        for (int k = lane; k < L; k += 32) {
            const double v = (double)x[lag + k];
            a += v * v;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0) {
            energy[row * n_lag + lag] = (float)a;
            if (lag == max_lag) sst[row] = (float)a;
        }
    }
}

__global__ void __launch_bounds__(IT_THREADS)
    intertp_cc_kernel(const float *__restrict__ wave, const float *__restrict__ weights,
                      const unsigned char *__restrict__ mask, const float *__restrict__ sst,
                      const float *__restrict__ energy, int T, int SC, int Lw, int max_lag,
                      float *__restrict__ out) {
    extern __shared__ __align__(16) float d_s[];  // Lw + 48 floats: the data waveform of (t, channel), zero padded
    const int L = Lw - 2 * max_lag, n_lag = 2 * max_lag + 1;
    const int t = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t2_first = blockIdx.x * IT_TB2 + warp * IT_PER_WARP;
    const int pad = ((Lw + 3) & ~3) + 44;

    float total[IT_PER_WARP];
    bool live[IT_PER_WARP];
#pragma unroll
    for (int j = 0; j < IT_PER_WARP; j++) {
        const int t2 = t2_first + j;
        total[j] = 0.f;
        live[j] = t2 < T && (mask == nullptr || mask[(size_t)t * T + t2] != 0);
    }

    for (int ch = 0; ch < SC; ch++) {
        const float w = weights[(size_t)t * SC + ch];
        if (w == 0.f) continue;  // CTA-uniform: zero-weight channels are skipped
        __syncthreads();
        const float *d = wave + ((size_t)t * SC + ch) * (size_t)Lw;
        for (int i = threadIdx.x; i < pad; i += IT_THREADS) d_s[i] = (i < Lw) ? d[i] : 0.f;
        __syncthreads();
        const float *en = energy + ((size_t)t * SC + ch) * n_lag;
#pragma unroll
        for (int j = 0; j < IT_PER_WARP; j++) {
            if (!live[j]) continue;  // warp-uniform
            const int t2 = t2_first + j;
            const float *x = wave + ((size_t)t2 * SC + ch) * (size_t)Lw + max_lag;
            const float ss = sst[(size_t)t2 * SC + ch];
            const float ss_root = sqrtf(ss);
            float best = -INFINITY;
            for (int lag0 = 0; lag0 < n_lag; lag0 += IT_LB) {
                float acc[IT_LB];
#pragma unroll
                for (int q = 0; q < IT_LB; q++) acc[q] = 0.f;
                for (int k = 4 * lane; k < L; k += 128) {
                    float xv[4];
#pragma unroll
                    for (int i = 0; i < 4; i++) xv[i] = (k + i < L) ? __ldg(x + k + i) : 0.f;
                    // d[k + lag0 + i + q], i < 4, q < 8: 11 floats starting at a multiple of 4 (16-byte aligned)
                    float win[12];
                    const float4 *d4 = reinterpret_cast<const float4 *>(d_s + k + lag0);
#pragma unroll
                    for (int v = 0; v < 3; v++) {
                        const float4 f = d4[v];
                        win[4 * v] = f.x; win[4 * v + 1] = f.y; win[4 * v + 2] = f.z; win[4 * v + 3] = f.w;
                    }
#pragma unroll
                    for (int q = 0; q < IT_LB; q++)
#pragma unroll
                        for (int i = 0; i < 4; i++) acc[q] = fmaf(xv[i], win[i + q], acc[q]);
                }
#pragma unroll
                for (int q = 0; q < IT_LB; q++) {
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], o);
                    if (lag0 + q < n_lag) {
                        const float e_l = en[lag0 + q];
                        const float v = ((double)ss * (double)e_l > (double)MF_GUARD) ? acc[q] / (ss_root * sqrtf(e_l)) : 0.f;
                        best = fmaxf(best, v);
                    }
                }
            }
            total[j] = fmaf(w, best, total[j]);
        }
    }
    if (lane == 0) {
#pragma unroll
        for (int j = 0; j < IT_PER_WARP; j++) {
            const int t2 = t2_first + j;
            if (t2 < T) out[(size_t)t * T + t2] = live[j] ? total[j] : 0.f;
        }
    }
}

}  // namespace bpmf

using namespace bpmf;

extern "C" size_t bpmf_intertemplate_workspace_bytes(size_t n_templates, size_t n_stations, size_t n_components,
                                                     size_t max_lag) {
    Workspace ws(nullptr, 0);
    const size_t rows = n_templates * n_stations * n_components;
    ws.take<float>(rows);
    ws.take<float>(rows * (2 * max_lag + 1));
    return ws.used;
}

extern "C" int bpmf_intertemplate_cc_dev(const float *d_waveforms, const float *d_station_weights,
                                         const unsigned char *d_pair_mask, size_t n_templates, size_t n_stations,
                                         size_t n_components, size_t n_samples, size_t max_lag, void *d_workspace,
                                         size_t workspace_bytes, float *d_cc, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const size_t T = n_templates, SC = n_stations * n_components, Lw = n_samples;
    if (T == 0) return BPMF_OK;
    BPMF_REQUIRE(SC > 0 && Lw > 2 * max_lag, "intertemplate_cc: waveforms of %zu samples cannot be cut by max_lag=%zu",
                 Lw, max_lag);
    BPMF_REQUIRE(T < (1u << 16) * (size_t)IT_TB2 && T <= 65535 && Lw < (1u << 24), "intertemplate_cc: too many templates");
    const size_t rows = T * SC;
    Workspace ws(d_workspace, workspace_bytes);
    float *sst = ws.take<float>(rows);
    float *energy = ws.take<float>(rows * (2 * max_lag + 1));
    BPMF_REQUIRE(ws.ok(), "intertemplate_cc: workspace too small (%zu < %zu)", workspace_bytes, ws.used);
    intertp_stats_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(d_waveforms, rows, (int)Lw, (int)max_lag, sst,
                                                                     energy);
    BPMF_LAUNCH_CHECK();
    const size_t smem = sizeof(float) * (((Lw + 3) & ~(size_t)3) + 48);
    BPMF_REQUIRE(smem <= 200 * 1024, "intertemplate_cc: waveforms too long for one shared-memory tile (%zu samples)", Lw);
    BPMF_CUDA(cudaFuncSetAttribute(intertp_cc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)((T + IT_TB2 - 1) / IT_TB2), (unsigned)T);
    {
        MainKernelTimer timer(st);
        intertp_cc_kernel<<<grid, IT_THREADS, smem, st>>>(d_waveforms, d_station_weights, d_pair_mask, sst, energy,
                                                          (int)T, (int)SC, (int)Lw, (int)max_lag, d_cc);
    }
    BPMF_LAUNCH_CHECK();
    return BPMF_OK;
}

extern "C" int bpmf_intertemplate_cc(const float *waveforms, const float *station_weights,
                                     const unsigned char *pair_mask, size_t n_templates, size_t n_stations,
                                     size_t n_components, size_t n_samples, size_t max_lag, float *cc) {
    const size_t T = n_templates, SC = n_stations * n_components;
    if (T == 0) return BPMF_OK;
    DevBuf w, sw, pm, ws, out;
    BPMF_CUDA(w.alloc(sizeof(float) * T * SC * n_samples));
    BPMF_CUDA(sw.alloc(sizeof(float) * T * SC));
    BPMF_CUDA(out.alloc(sizeof(float) * T * T));
    const size_t wsb = bpmf_intertemplate_workspace_bytes(T, n_stations, n_components, max_lag);
    BPMF_CUDA(ws.alloc(wsb));
    cudaStream_t st = nullptr;
    BPMF_CUDA(cudaMemcpyAsync(w.p, waveforms, sizeof(float) * T * SC * n_samples, cudaMemcpyHostToDevice, st));
    BPMF_CUDA(cudaMemcpyAsync(sw.p, station_weights, sizeof(float) * T * SC, cudaMemcpyHostToDevice, st));
    if (pair_mask) {
        BPMF_CUDA(pm.alloc(T * T));
        BPMF_CUDA(cudaMemcpyAsync(pm.p, pair_mask, T * T, cudaMemcpyHostToDevice, st));
    }
    const int rc = bpmf_intertemplate_cc_dev(w.as<float>(), sw.as<float>(), pair_mask ? pm.as<unsigned char>() : nullptr,
                                             T, n_stations, n_components, n_samples, max_lag, ws.p, wsb,
                                             out.as<float>(), st);
    if (rc != BPMF_OK) return rc;
    BPMF_CUDA(cudaMemcpyAsync(cc, out.p, sizeof(float) * T * T, cudaMemcpyDeviceToHost, st));
    BPMF_CUDA(cudaStreamSynchronize(st));
    return BPMF_OK;
}
