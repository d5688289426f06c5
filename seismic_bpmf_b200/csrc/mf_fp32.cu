// mf_fp32.cu -- direct-form matched filter on FP32 CUDA cores (BPMF_MF_FP32).
//
// Replaces the native kernel behind fast_matched_filter.matched_filter as BPMF calls it
// (BPMF/similarity_search.py:526-533, BPMF/dataset.py:4818-4827); semantics: SURVEY.md A.1.
//
// One CTA = one template x one tile of 4096 consecutive outputs, looping over the weighted
// channels of that template.  Per channel the trace segment (tile + L samples, shifted by the
// channel's moveout) and the template are staged in shared memory; each thread owns 16
// consecutive outputs and slides a 20-register window over the segment, so one LDS.128 of trace
// and one broadcast LDS.128 of template feed 64 FMAs (FMA-issue-bound, not LDS-bound).  The
// segment is stored with 4 floats of padding per 16 so the per-thread 64-byte stride is
// bank-conflict free.  Roofline: FP32 FMA pipe, 2*L flop per (template, channel, output).
#include "common.cuh"

namespace bpmf {

constexpr int F_THREADS = 256;
constexpr int F_R = 16;
constexpr int F_TI = F_THREADS * F_R;  // 4096 outputs per tile (step == 1 kernel)

__host__ __device__ __forceinline__ int pad_idx(int i) { return i + ((i >> 4) << 2); }

template <bool FULL, bool NETSUM>
__global__ void __launch_bounds__(F_THREADS, 2)
    mf_fp32_step1_kernel(const float *__restrict__ templates, const int *__restrict__ moveouts,
                         const float *__restrict__ weights, const float *__restrict__ data,
                         const float *__restrict__ sst, const float *__restrict__ sumt,
                         const int *__restrict__ lo, const int *__restrict__ hi,
                         const float *__restrict__ energy, const float *__restrict__ mean,
                         float *__restrict__ cc, size_t cc_pitch, long long N, int L,
                         long long n_corr, int SC, long long n_win) {
    extern __shared__ __align__(16) float smem[];
    const int nkb_ceil = (L + 3) >> 2;
    const int span_read = F_TI + 4 * nkb_ceil;  // samples staged per channel
    float *d_s = smem;
    float *t_s = smem + pad_idx(span_read + 16);

    const int tid = threadIdx.x;
    const size_t t = blockIdx.y;
    const long long n0 = (long long)blockIdx.x * F_TI;
    const long long lo_t = lo[t], hi_t = hi[t];
    const long long n_first = n0 + (long long)F_R * tid;

    float total[F_R];
#pragma unroll
    for (int r = 0; r < F_R; r++) total[r] = 0.f;

    const bool tile_live = (hi_t >= lo_t) && (n0 <= hi_t) && (n0 + F_TI - 1 >= lo_t);
    if (tile_live) {
        for (int ch = 0; ch < SC; ch++) {
            const float w = weights[t * SC + ch];
            if (w == 0.f) continue;  // CTA-uniform
            const long long mv = moveouts[t * SC + ch];
            const long long j0 = n0 + mv;
            const float *dch = data + (size_t)ch * N;
            for (int i = tid; i < span_read; i += F_THREADS) {
                const long long g = j0 + i;
                d_s[pad_idx(i)] = (g >= 0 && g < N) ? __ldg(dch + g) : 0.f;
            }
            const float *tp = templates + (t * SC + ch) * (size_t)L;
            for (int k = tid; k < 4 * nkb_ceil; k += F_THREADS) t_s[k] = (k < L) ? __ldg(tp + k) : 0.f;
            __syncthreads();

            const float4 *d4 = reinterpret_cast<const float4 *>(d_s);
            const float4 *t4 = reinterpret_cast<const float4 *>(t_s);
            float win[20];
            float acc[F_R];
#pragma unroll
            for (int r = 0; r < F_R; r++) acc[r] = 0.f;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const float4 v = d4[5 * tid + q];
                win[4 * q + 0] = v.x; win[4 * q + 1] = v.y; win[4 * q + 2] = v.z; win[4 * q + 3] = v.w;
            }
            const int nkb_full = L >> 2;
            int kb = 0;
            // main loop: groups of 5 k-blocks so that the 20-slot window rotation is static
            for (; kb + 5 <= nkb_full; kb += 5) {
#pragma unroll
                for (int b = 0; b < 5; b++) {
                    const int kq = kb + b;
                    const float4 nv = d4[5 * tid + 5 + kq + (kq >> 2)];
                    win[(4 * b + 16) % 20] = nv.x; win[(4 * b + 17) % 20] = nv.y;
                    win[(4 * b + 18) % 20] = nv.z; win[(4 * b + 19) % 20] = nv.w;
                    const float4 tv = t4[kq];
#pragma unroll
                    for (int r = 0; r < F_R; r++) {
                        acc[r] = fmaf(tv.x, win[(4 * b + r + 0) % 20], acc[r]);
                        acc[r] = fmaf(tv.y, win[(4 * b + r + 1) % 20], acc[r]);
                        acc[r] = fmaf(tv.z, win[(4 * b + r + 2) % 20], acc[r]);
                        acc[r] = fmaf(tv.w, win[(4 * b + r + 3) % 20], acc[r]);
                    }
                }
            }
            // tail: < 5 full blocks plus an optional partial block (guards are CTA-uniform)
            if (kb < nkb_ceil) {
#pragma unroll
                for (int b = 0; b < 5; b++) {
                    const int kq = kb + b;
                    if (kq < nkb_ceil) {
                        const float4 nv = d4[5 * tid + 5 + kq + (kq >> 2)];
                        win[(4 * b + 16) % 20] = nv.x; win[(4 * b + 17) % 20] = nv.y;
                        win[(4 * b + 18) % 20] = nv.z; win[(4 * b + 19) % 20] = nv.w;
                        const float4 tv = t4[kq];
                        const int rem = L - 4 * kq;  // >= 1
#pragma unroll
                        for (int r = 0; r < F_R; r++) {
                            acc[r] = fmaf(tv.x, win[(4 * b + r + 0) % 20], acc[r]);
                            if (rem > 1) acc[r] = fmaf(tv.y, win[(4 * b + r + 1) % 20], acc[r]);
                            if (rem > 2) acc[r] = fmaf(tv.z, win[(4 * b + r + 2) % 20], acc[r]);
                            if (rem > 3) acc[r] = fmaf(tv.w, win[(4 * b + r + 3) % 20], acc[r]);
                        }
                    }
                }
            }

            // normalise + accumulate
            const float sst_v = sst[t * SC + ch];
            const float sst_root = sqrtf(sst_v);
            const float sumt_v = FULL ? sumt[t * SC + ch] : 0.f;
            const float *ech = energy + (size_t)ch * n_win;
            const float *mch = FULL ? mean + (size_t)ch * n_win : nullptr;
#pragma unroll
            for (int r = 0; r < F_R; r++) {
                const long long n = n_first + r;
                if (n >= lo_t && n <= hi_t) {
                    const long long j = n + mv;
                    float num = acc[r];
                    if (FULL) num -= mch[j] * sumt_v;
                    // guard on the float64 product, quotient by the two roots: sst * E itself overflows
                    // float32 for raw-count amplitudes (~1e13 transients)
                    const float e_j = ech[j];
                    const float v = ((double)sst_v * (double)e_j > (double)MF_GUARD) ? num / (sst_root * sqrtf(e_j)) : 0.f;
                    if (NETSUM)
                        total[r] = fmaf(w, v, total[r]);
                    else
                        cc[((size_t)t * n_corr + (size_t)n) * SC + ch] = v;
                }
            }
            __syncthreads();
        }
    }
    if (NETSUM) {
        float *row = cc + t * cc_pitch;
#pragma unroll
        for (int r = 0; r < F_R; r++) {
            const long long n = n_first + r;
            if (n < n_corr) row[n] = total[r];
        }
    }
}

// ---------------------------------------------------------------- generic step (>= 1)
template <int R, bool FULL, bool NETSUM>
__global__ void __launch_bounds__(F_THREADS)
    mf_fp32_generic_kernel(const float *__restrict__ templates, const int *__restrict__ moveouts,
                           const float *__restrict__ weights, const float *__restrict__ data,
                           const float *__restrict__ sst, const float *__restrict__ sumt,
                           const int *__restrict__ lo, const int *__restrict__ hi,
                           const float *__restrict__ energy, const float *__restrict__ mean,
                           float *__restrict__ cc, size_t cc_pitch, long long N, int L,
                           long long n_corr, int SC, long long n_win, long long step) {
    extern __shared__ __align__(16) float smem[];
    constexpr int TI = F_THREADS * R;
    const int span = (int)((TI - 1) * step) + L;
    float *d_s = smem;
    float *t_s = smem + span;
    const int tid = threadIdx.x;
    const size_t t = blockIdx.y;
    const long long n0 = (long long)blockIdx.x * TI;
    const long long lo_t = lo[t], hi_t = hi[t];
    float total[R];
#pragma unroll
    for (int r = 0; r < R; r++) total[r] = 0.f;
    const bool tile_live = (hi_t >= lo_t) && (n0 <= hi_t) && (n0 + TI - 1 >= lo_t);
    if (tile_live) {
        for (int ch = 0; ch < SC; ch++) {
            const float w = weights[t * SC + ch];
            if (w == 0.f) continue;
            const long long mv = moveouts[t * SC + ch];
            const long long j0 = n0 * step + mv;
            const float *dch = data + (size_t)ch * N;
            for (int i = tid; i < span; i += F_THREADS) {
                const long long g = j0 + i;
                d_s[i] = (g >= 0 && g < N) ? __ldg(dch + g) : 0.f;
            }
            const float *tp = templates + (t * SC + ch) * (size_t)L;
            for (int k = tid; k < L; k += F_THREADS) t_s[k] = __ldg(tp + k);
            __syncthreads();
            float acc[R];
#pragma unroll
            for (int r = 0; r < R; r++) acc[r] = 0.f;
            for (int k = 0; k < L; k++) {
                const float tk = t_s[k];
#pragma unroll
                for (int r = 0; r < R; r++)
                    acc[r] = fmaf(tk, d_s[(long long)(tid + F_THREADS * r) * step + k], acc[r]);
            }
            const float sst_v = sst[t * SC + ch];
            const float sst_root = sqrtf(sst_v);
            const float sumt_v = FULL ? sumt[t * SC + ch] : 0.f;
            const float *ech = energy + (size_t)ch * n_win;
            const float *mch = FULL ? mean + (size_t)ch * n_win : nullptr;
#pragma unroll
            for (int r = 0; r < R; r++) {
                const long long n = n0 + tid + F_THREADS * r;
                if (n >= lo_t && n <= hi_t) {
                    const long long j = n * step + mv;
                    float num = acc[r];
                    if (FULL) num -= mch[j] * sumt_v;
                    // guard on the float64 product, quotient by the two roots: sst * E itself overflows
                    // float32 for raw-count amplitudes (~1e13 transients)
                    const float e_j = ech[j];
                    const float v = ((double)sst_v * (double)e_j > (double)MF_GUARD) ? num / (sst_root * sqrtf(e_j)) : 0.f;
                    if (NETSUM)
                        total[r] = fmaf(w, v, total[r]);
                    else
                        cc[((size_t)t * n_corr + (size_t)n) * SC + ch] = v;
                }
            }
            __syncthreads();
        }
    }
    if (NETSUM) {
        float *row = cc + t * cc_pitch;
#pragma unroll
        for (int r = 0; r < R; r++) {
            const long long n = n0 + tid + F_THREADS * r;
            if (n < n_corr) row[n] = total[r];
        }
    }
}

template <typename K>
static int set_smem(K kernel, size_t bytes) {
    BPMF_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return BPMF_OK;
}

#define MF_ARGS                                                                                 \
    d_templates, d_moveouts, d_weights, d_data, meta.sst, meta.sumt, meta.lo, meta.hi, d_energy, \
        d_mean, d_cc, cc_pitch, (long long)sh.N, (int)sh.L, (long long)sh.n_corr, (int)sh.SC(),  \
        (long long)n_win

int mf_fp32_run(const MfShape &sh, const float *d_templates, const int *d_moveouts,
                const float *d_weights, const float *d_data, MfMeta meta, const float *d_energy,
                const float *d_mean, float *d_cc, size_t cc_pitch, cudaStream_t st) {
    const size_t n_win = sh.N - sh.L + 1;
    const bool full = sh.normalize != 0, netsum = sh.network_sum != 0;
    const size_t SMEM_MAX = 220 * 1024;
    int rc;
    MainKernelTimer timer(st);
    if (sh.step == 1) {
        const int nkb = (int)((sh.L + 3) / 4);
        const size_t smem = sizeof(float) * ((size_t)pad_idx(F_TI + 4 * nkb + 16) + 4 * (size_t)nkb);
        if (smem <= SMEM_MAX) {
            dim3 grid((unsigned)((sh.n_corr + F_TI - 1) / F_TI), (unsigned)sh.T);
#define LAUNCH1(FU, NS)                                                                      \
    do {                                                                                     \
        if ((rc = set_smem(mf_fp32_step1_kernel<FU, NS>, smem)) != BPMF_OK) return rc;       \
        mf_fp32_step1_kernel<FU, NS><<<grid, F_THREADS, smem, st>>>(MF_ARGS);                \
    } while (0)
            if (full && netsum) LAUNCH1(true, true);
            else if (full) LAUNCH1(true, false);
            else if (netsum) LAUNCH1(false, true);
            else LAUNCH1(false, false);
#undef LAUNCH1
            BPMF_LAUNCH_CHECK();
            return BPMF_OK;
        }
    }
    // generic path
    const size_t span4 = (size_t)(F_THREADS * 4 - 1) * sh.step + sh.L;
    const size_t span1 = (size_t)(F_THREADS - 1) * sh.step + sh.L;
    const bool r4 = sizeof(float) * (span4 + sh.L) <= SMEM_MAX;
    const size_t smem = sizeof(float) * ((r4 ? span4 : span1) + sh.L);
    BPMF_REQUIRE(smem <= SMEM_MAX, "matched filter: step=%zu with L=%zu exceeds the shared-memory tile",
                 sh.step, sh.L);
    const int TI = F_THREADS * (r4 ? 4 : 1);
    dim3 grid((unsigned)((sh.n_corr + TI - 1) / TI), (unsigned)sh.T);
#define LAUNCHG(R, FU, NS)                                                                       \
    do {                                                                                         \
        if ((rc = set_smem(mf_fp32_generic_kernel<R, FU, NS>, smem)) != BPMF_OK) return rc;      \
        mf_fp32_generic_kernel<R, FU, NS><<<grid, F_THREADS, smem, st>>>(MF_ARGS, (long long)sh.step); \
    } while (0)
    if (r4) {
        if (full && netsum) LAUNCHG(4, true, true);
        else if (full) LAUNCHG(4, true, false);
        else if (netsum) LAUNCHG(4, false, true);
        else LAUNCHG(4, false, false);
    } else {
        if (full && netsum) LAUNCHG(1, true, true);
        else if (full) LAUNCHG(1, true, false);
        else if (netsum) LAUNCHG(1, false, true);
        else LAUNCHG(1, false, false);
    }
#undef LAUNCHG
    BPMF_LAUNCH_CHECK();
    return BPMF_OK;
}

}  // namespace bpmf
