// mf_prepass.cu -- matched-filter pre-pass kernels shared by the fp32 and tensor paths:
//   * per-(template, channel) sum of squares / sum           (fast_matched_filter computes
//     sum_square_templates on the host before its native call [UPSTREAM-UNVERIFIED])
//   * per-template valid output range from weights + moveouts
//   * per-channel sliding window energy  E[ch][j] = sum_{k<L} d[ch][j+k]^2, j in [0, N-L]
//     accumulated in float64 segment-local sums, free of cancellation (SURVEY section 7 "Sliding energy in FP32").
// All HBM-bound streaming kernels: reads 4*S*C*N, writes 4*S*C*(N-L+1) (x2 for normalize=full).
#include "common.cuh"

namespace bpmf {

// ---------------------------------------------------------------- template statistics
__global__ void mf_template_stats_kernel(const float *__restrict__ templates, size_t n_rows,
                                         size_t L, float *__restrict__ sst,
                                         float *__restrict__ sumt) {
    const size_t row = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= n_rows) return;
    const float *p = templates + row * L;
    double a = 0.0, b = 0.0;
    for (size_t k = lane; k < L; k += 32) {
        const double v = (double)p[k];
        a += v * v;
        b += v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if (lane == 0) {
        sst[row] = (float)a;
        sumt[row] = (float)b;
    }
}

// ---------------------------------------------------------------- per-template output bounds
__global__ void mf_bounds_kernel(const int *__restrict__ moveouts, const float *__restrict__ weights,
                                 size_t T, size_t SC, long long N, long long L, long long step,
                                 long long n_corr, int *__restrict__ lo, int *__restrict__ hi) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    long long mn = 0, mx = 0;
    bool any = false;
    for (size_t ch = 0; ch < SC; ch++) {
        if (weights[t * SC + ch] == 0.f) continue;
        const long long m = moveouts[t * SC + ch];
        if (!any) { mn = mx = m; any = true; }
        mn = m < mn ? m : mn;
        mx = m > mx ? m : mx;
    }
    long long n_lo = 0, n_hi = -1;
    if (any) {
        if (mn < 0) n_lo = (-mn + step - 1) / step;
        const long long last = N - L - mx;  // last valid window start of the slowest channel
        if (last >= 0) {
            n_hi = last / step;
            if (n_hi > n_corr - 1) n_hi = n_corr - 1;
        }
    }
    lo[t] = (int)n_lo;
    hi[t] = (int)n_hi;
}

// ---------------------------------------------------------------- sliding window energy
// E[j] = sum_{k<L} d[j+k]^2 without any cancellation: the block's samples are cut into segments of L
// samples; with SUF[i] = sum from i to the end of i's segment and PRE[i] = sum from the start of i's
// segment to i, a window starting at j = m*L + r is the tail of segment m plus the first r samples of
// segment m+1:  E[j] = SUF[j] + (r ? PRE[j+L-1] : 0).  Only sums of at most L same-sign terms are ever
// formed (float64), so a 1e7 x noise transient does not leak into the energy of the quiet windows after
// it, which a running prefix-sum difference (what mf_oracle.c and upstream do) cannot guarantee.
constexpr int EN_THREADS = 256;
constexpr int EN_R = 16;
constexpr int EN_BJ = EN_THREADS * EN_R;  // window starts per block

// Segmented scan over the threads of the block.  Every thread contributes (tail, flag): the sum of its
// run after the last segment start inside the run, and whether the run contains a segment start.
// Returns the carry: the sum of the elements between the closest preceding segment start and the
// thread's own run.
__device__ __forceinline__ double block_segmented_carry(double tail, bool flag, double *warp_val, int *warp_flag) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    double s = tail;
    int f = flag ? 1 : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double us = __shfl_up_sync(0xffffffffu, s, o);
        const int uf = __shfl_up_sync(0xffffffffu, f, o);
        if (lane >= o) {
            if (!f) s += us;
            f |= uf;
        }
    }
    if (lane == 31) { warp_val[wid] = s; warp_flag[wid] = f; }
    double ex_s = __shfl_up_sync(0xffffffffu, s, 1);
    int ex_f = __shfl_up_sync(0xffffffffu, f, 1);
    if (lane == 0) { ex_s = 0.0; ex_f = 0; }
    __syncthreads();
    double ws = 0.0;  // aggregate of the warps before this one
#pragma unroll
    for (int w = 0; w < EN_THREADS / 32; w++) {
        if (w < wid) ws = warp_flag[w] ? warp_val[w] : ws + warp_val[w];
    }
    __syncthreads();
    return ex_f ? ex_s : ws + ex_s;
}

// REV = false: out[i] = sum of x from the start of i's segment to i (segments of L samples from 0);
// REV = true:  out[i] = sum of x from i to the end of i's segment (or of the span); x = d^2 or d.
template <bool REV, bool SQ>
__device__ __forceinline__ void segment_sums(const float *__restrict__ d, size_t span, size_t L,
                                             double *__restrict__ out, double *warp_val, int *warp_flag) {
    const size_t run = (span + EN_THREADS - 1) / EN_THREADS;
    const size_t k = REV ? (size_t)(EN_THREADS - 1 - threadIdx.x) : (size_t)threadIdx.x;
    const size_t a = (k * run < span) ? k * run : span;
    const size_t b = (a + run < span) ? a + run : span;
    double tail = 0.0;
    bool flag = false;
    if (!REV) {
        size_t pos = a % L;
        for (size_t i = a; i < b; i++) {
            if (pos == 0) { tail = 0.0; flag = true; }
            if (++pos == L) pos = 0;
            const double v = (double)d[i];
            tail += SQ ? v * v : v;
        }
    } else {
        size_t rem = b % L;  // (i + 1) % L for i = b - 1
        for (size_t i = b; i-- > a;) {
            if (rem == 0 || i == span - 1) { tail = 0.0; flag = true; }
            if (rem == 0) rem = L;
            rem--;
            const double v = (double)d[i];
            tail += SQ ? v * v : v;
        }
    }
    double r = block_segmented_carry(tail, flag, warp_val, warp_flag);
    if (!REV) {
        size_t pos = a % L;
        for (size_t i = a; i < b; i++) {
            if (pos == 0) r = 0.0;
            if (++pos == L) pos = 0;
            const double v = (double)d[i];
            r += SQ ? v * v : v;
            out[i] = r;
        }
    } else {
        size_t rem = b % L;
        for (size_t i = b; i-- > a;) {
            if (rem == 0 || i == span - 1) r = 0.0;
            if (rem == 0) rem = L;
            rem--;
            const double v = (double)d[i];
            r += SQ ? v * v : v;
            out[i] = r;
        }
    }
}

// RSQRT: write 1/sqrt(E) (0 where E <= 0) instead of E -- the per-column factor of the tensor path.
// phase8: write phase-major, out[ch][j % 8][j / 8] with rows of out_pitch elements.
template <bool FULL, bool RSQRT>
__global__ void __launch_bounds__(EN_THREADS)
    mf_energy_kernel(const float *__restrict__ data, size_t N, size_t L, size_t n_win, size_t out_pitch,
                     int phase8, float *__restrict__ energy, float *__restrict__ mean) {
    extern __shared__ double sm[];
    __shared__ double warp_val[EN_THREADS / 32];
    __shared__ int warp_flag[EN_THREADS / 32];
    const size_t ch = blockIdx.y;
    const size_t j0 = (size_t)blockIdx.x * EN_BJ;
    if (j0 >= n_win) return;
    const size_t nj = (n_win - j0 < (size_t)EN_BJ) ? (n_win - j0) : (size_t)EN_BJ;
    const size_t span = nj + L - 1;  // data samples [j0, j0 + span)
    const float *d = data + ch * N + j0;
    double *P2 = sm;                         // span sums of d^2
    double *P1 = sm + (FULL ? span : 0);     // span sums of d

    // tails of the segments (kept in registers), then the heads of the next segments in the same arrays
    double suf2[EN_R], suf1[EN_R];
    segment_sums<true, true>(d, span, L, P2, warp_val, warp_flag);
    if (FULL) segment_sums<true, false>(d, span, L, P1, warp_val, warp_flag);
    __syncthreads();
#pragma unroll
    for (int m = 0; m < EN_R; m++) {
        const size_t j = (size_t)threadIdx.x + (size_t)EN_THREADS * m;
        suf2[m] = (j < nj) ? P2[j] : 0.0;
        suf1[m] = (FULL && j < nj) ? P1[j] : 0.0;
    }
    __syncthreads();
    segment_sums<false, true>(d, span, L, P2, warp_val, warp_flag);
    if (FULL) segment_sums<false, false>(d, span, L, P1, warp_val, warp_flag);
    __syncthreads();

    float *e = energy + (phase8 ? ch * 8 * out_pitch : ch * out_pitch + j0);
    float *mo = FULL ? mean + ch * out_pitch + j0 : nullptr;
#pragma unroll
    for (int m = 0; m < EN_R; m++) {
        const size_t j = (size_t)threadIdx.x + (size_t)EN_THREADS * m;
        if (j >= nj) continue;
        const bool whole = (j % L) == 0;  // the window is exactly one segment
        double ss = suf2[m] + (whole ? 0.0 : P2[j + L - 1]);
        if (FULL) {
            const double mu = (suf1[m] + (whole ? 0.0 : P1[j + L - 1])) / (double)L;
            ss -= (double)L * mu * mu;
            mo[j] = (float)mu;
        }
        const size_t jg = j0 + j;
        const size_t o = phase8 ? (jg & 7) * out_pitch + (jg >> 3) : j;
        if (RSQRT) e[o] = (ss > 0.0) ? (float)(1.0 / sqrt(ss)) : 0.f;
        else e[o] = (float)(ss > 0.0 ? ss : 0.0);
    }
}

int mf_energy(const MfShape &sh, const float *d_data, float *d_energy, float *d_mean, size_t out_pitch,
              bool rsqrt_mode, bool phase8, cudaStream_t st) {
    const size_t n_win = sh.N - sh.L + 1;
    const bool full = sh.normalize != 0;
    const size_t span_max = (size_t)EN_BJ + sh.L;
    const size_t smem = sizeof(double) * span_max * (full ? 2 : 1);
    BPMF_REQUIRE(smem <= 200 * 1024, "template too long for the energy pre-pass (L=%zu)", sh.L);
    dim3 grid((unsigned)((n_win + EN_BJ - 1) / EN_BJ), (unsigned)sh.SC());
#define EN_LAUNCH(FU, RS)                                                                          \
    do {                                                                                           \
        BPMF_CUDA(cudaFuncSetAttribute(mf_energy_kernel<FU, RS>,                                   \
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   \
        mf_energy_kernel<FU, RS><<<grid, EN_THREADS, smem, st>>>(d_data, sh.N, sh.L, n_win,        \
                                                                 out_pitch, phase8 ? 1 : 0,        \
                                                                 d_energy, d_mean);                \
    } while (0)
    if (full && rsqrt_mode) EN_LAUNCH(true, true);
    else if (full) EN_LAUNCH(true, false);
    else if (rsqrt_mode) EN_LAUNCH(false, true);
    else EN_LAUNCH(false, false);
#undef EN_LAUNCH
    BPMF_LAUNCH_CHECK();
    return BPMF_OK;
}

int mf_prepass(const MfShape &sh, const float *d_templates, const int *d_moveouts,
               const float *d_weights, const float *d_data, MfMeta meta, float *d_energy,
               float *d_mean, cudaStream_t st) {
    const size_t rows = sh.T * sh.SC();
    {
        const int warps = 8;
        mf_template_stats_kernel<<<(unsigned)((rows + warps - 1) / warps), warps * 32, 0, st>>>(
            d_templates, rows, sh.L, meta.sst, meta.sumt);
        BPMF_LAUNCH_CHECK();
    }
    {
        mf_bounds_kernel<<<(unsigned)((sh.T + 127) / 128), 128, 0, st>>>(
            d_moveouts, d_weights, sh.T, sh.SC(), (long long)sh.N, (long long)sh.L,
            (long long)sh.step, (long long)sh.n_corr, meta.lo, meta.hi);
        BPMF_LAUNCH_CHECK();
    }
    if (d_energy) {
        const size_t n_win = sh.N - sh.L + 1;
        int rc = mf_energy(sh, d_data, d_energy, d_mean, n_win, false, false, st);
        if (rc != BPMF_OK) return rc;
    }
    return BPMF_OK;
}

}  // namespace bpmf
