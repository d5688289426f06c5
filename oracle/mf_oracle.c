/*
 * oracle/mf_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (C99 + OpenMP) of the matched-filter CC-sum that BPMF obtains from the
 * un-vendored pip dependency `FastMatchedFilter` (pyproject.toml:29 of the reference, no version
 * pin, absent from /root/reference and from this image).  PARITY UNPINNED at the kernel boundary:
 * there is no reference source, golden vector or test for this path; the semantics below follow
 * SURVEY.md Appendix A.1, whose pinned lines are the reference's own call sites:
 *   - argument order / shapes      BPMF/similarity_search.py:505-506,526-533
 *   - int32 moveouts (may be < 0)  BPMF/dataset.py:4122-4125,3452-3475
 *   - additivity over stations     BPMF/similarity_search.py:517-538
 *   - exact 0.0 on dead windows    BPMF/similarity_search.py:615 ; BPMF/libc.c:561,609
 *   - network_sum=False layout     BPMF/dataset.py:4818-4830  -> (T, n_corr, S, C)
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * call this file.  The product path (seismic_bpmf_b200/csrc) never links or loads it.
 *
 * Semantics (A.1):
 *   n_corr = (N - L)/step + 1
 *   cc_sum[t, n] = sum_{s,c : w != 0} w[t,s,c] * CC(T[t,s,c,:], D[s,c, n*step + mv[t,s,c] : +L])
 *   CC = <T,D> / sqrt(sum T^2 * sum D^2)   ("short" normalisation; "full" removes the window mean)
 *   a channel contributes exactly 0 when sum T^2 * sum D^2 <= 1e-6   (stability guard)
 *   an output index stays 0 when ANY weighted channel window leaves [0, N)
 */
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <omp.h>

#define MF_STABILITY_GUARD 0.000001f

/* double-precision prefix sums of d and d^2 for one channel: out has N+1 entries */
static void prefix_sums(const float *d, size_t n, double *csum, double *csum2) {
    double a = 0.0, b = 0.0;
    csum[0] = 0.0;
    csum2[0] = 0.0;
    for (size_t i = 0; i < n; i++) {
        a += (double)d[i];
        b += (double)d[i] * (double)d[i];
        csum[i + 1] = a;
        csum2[i + 1] = b;
    }
}

static inline float dotf(const float *a, const float *b, size_t n) {
    float acc = 0.f;
#pragma omp simd reduction(+ : acc)
    for (size_t k = 0; k < n; k++) acc += a[k] * b[k];
    return acc;
}

/*
 * network_sum != 0: out is (T, n_corr), weighted sum over channels.
 * network_sum == 0: out is (T, n_corr, S, C), UNWEIGHTED per-channel CC.
 * normalize: 0 = "short", 1 = "full".
 * out must be zero-initialised by the caller.  Returns 0, or -1 on allocation failure.
 */
int oracle_matched_filter(const float *templates, const int *moveouts, const float *weights,
                          const float *data, size_t step, size_t L, size_t N, size_t T, size_t S,
                          size_t C, size_t n_corr, int normalize, int network_sum,
                          int num_threads, float *out) {
    const size_t SC = S * C;
    if (num_threads > 0) omp_set_num_threads(num_threads);
    if (N < L || step == 0) return 0;

    double *csum = (double *)malloc(sizeof(double) * SC * (N + 1));
    double *csum2 = (double *)malloc(sizeof(double) * SC * (N + 1));
    float *sst = (float *)malloc(sizeof(float) * T * SC);
    float *sumt = (float *)malloc(sizeof(float) * T * SC);
    if (!csum || !csum2 || !sst || !sumt) {
        free(csum); free(csum2); free(sst); free(sumt);
        return -1;
    }
#pragma omp parallel for schedule(dynamic)
    for (size_t ch = 0; ch < SC; ch++)
        prefix_sums(data + ch * N, N, csum + ch * (N + 1), csum2 + ch * (N + 1));
#pragma omp parallel for
    for (size_t q = 0; q < T * SC; q++) {
        double a = 0.0, b = 0.0;
        for (size_t k = 0; k < L; k++) {
            double v = (double)templates[q * L + k];
            a += v * v;
            b += v;
        }
        sst[q] = (float)a;
        sumt[q] = (float)b;
    }

    for (size_t t = 0; t < T; t++) {
        /* bounds over weighted channels */
        long min_mv = 0, max_mv = 0;
        int any = 0;
        for (size_t ch = 0; ch < SC; ch++) {
            if (weights[t * SC + ch] == 0.f) continue;
            long m = moveouts[t * SC + ch];
            if (!any) { min_mv = max_mv = m; any = 1; }
            if (m < min_mv) min_mv = m;
            if (m > max_mv) max_mv = m;
        }
        if (!any) continue;
        /* first / last valid output index n (data sample n*step) */
        long n_lo = 0;
        if (min_mv < 0) n_lo = (long)((-min_mv + (long)step - 1) / (long)step);
        long last_start = (long)N - (long)L - max_mv; /* inclusive */
        if (last_start < 0) continue;
        long n_hi = last_start / (long)step; /* inclusive */
        if (n_hi > (long)n_corr - 1) n_hi = (long)n_corr - 1;

#pragma omp parallel for schedule(static)
        for (long n = n_lo; n <= n_hi; n++) {
            float acc = 0.f;
            for (size_t ch = 0; ch < SC; ch++) {
                const float w = weights[t * SC + ch];
                if (w == 0.f) continue;
                const long j = n * (long)step + moveouts[t * SC + ch];
                const float *tp = templates + (t * SC + ch) * L;
                const float *dp = data + ch * N + j;
                float num = dotf(tp, dp, L);
                const double *c1 = csum + ch * (N + 1);
                const double *c2 = csum2 + ch * (N + 1);
                double ssd = c2[j + L] - c2[j];
                if (normalize) {
                    double mean = (c1[j + L] - c1[j]) / (double)L;
                    num -= (float)(mean * (double)sumt[t * SC + ch]);
                    ssd -= (double)L * mean * mean;
                }
                float den = sst[t * SC + ch] * (float)ssd;
                float cc = 0.f;
                if (den > MF_STABILITY_GUARD) cc = num / sqrtf(den);
                if (network_sum)
                    acc += w * cc;
                else
                    out[((t * n_corr) + (size_t)n) * SC + ch] = cc;
            }
            if (network_sum) out[t * n_corr + (size_t)n] = acc;
        }
    }
    free(csum); free(csum2); free(sst); free(sumt);
    return 0;
}

/*
 * Same semantics (step 1, "short" normalisation, network sum), evaluated only on selected runs of output
 * indices: out[t, w, x] = cc_sum[t, win_start[w] + x] for x < win_len (0 where the index is outside
 * [0, n_corr) or outside the template's valid range).  Used by the parity checks at full day length, where
 * the whole (T, n_corr) slab is hours of CPU work: the window energies are accumulated directly in double
 * precision (first window of a run: L terms; then one term in, one term out), no O(S*C*N) prefix arrays.
 */
int oracle_matched_filter_windows(const float *templates, const int *moveouts, const float *weights,
                                  const float *data, size_t L, size_t N, size_t T, size_t S, size_t C,
                                  size_t n_win, const long long *win_start, size_t win_len, int num_threads,
                                  float *out) {
    const size_t SC = S * C;
    if (num_threads > 0) omp_set_num_threads(num_threads);
    if (N < L) return 0;
    const long n_corr = (long)(N - L + 1);
    const size_t SUB = 32768; /* runs are cut into pieces so that a few long runs still use every core */
    const size_t n_sub = (win_len + SUB - 1) / SUB;
    float *sst = (float *)malloc(sizeof(float) * T * SC);
    if (!sst) return -1;
#pragma omp parallel for
    for (size_t q = 0; q < T * SC; q++) {
        double a = 0.0;
        for (size_t k = 0; k < L; k++) {
            double v = (double)templates[q * L + k];
            a += v * v;
        }
        sst[q] = (float)a;
    }
    memset(out, 0, sizeof(float) * T * n_win * win_len);
#pragma omp parallel for schedule(dynamic) collapse(3)
    for (size_t t = 0; t < T; t++) {
        for (size_t w = 0; w < n_win; w++) {
            for (size_t sb = 0; sb < n_sub; sb++) {
                long min_mv = 0, max_mv = 0;
                int any = 0;
                for (size_t ch = 0; ch < SC; ch++) {
                    if (weights[t * SC + ch] == 0.f) continue;
                    long m = moveouts[t * SC + ch];
                    if (!any) { min_mv = max_mv = m; any = 1; }
                    if (m < min_mv) min_mv = m;
                    if (m > max_mv) max_mv = m;
                }
                if (!any) continue;
                long n_lo = min_mv < 0 ? -min_mv : 0;
                long n_hi = (long)N - (long)L - max_mv; /* inclusive, as oracle_matched_filter */
                if (n_hi > n_corr - 1) n_hi = n_corr - 1;
                long a = (long)win_start[w] + (long)(sb * SUB);
                long b = a + (long)((win_len - sb * SUB < SUB) ? win_len - sb * SUB : SUB) - 1;
                float *o = out + (t * n_win + w) * win_len + sb * SUB - a; /* o[n] for n in [a, b] */
                if (a < n_lo) a = n_lo;
                if (b > n_hi) b = n_hi;
                if (b < a) continue;
                for (size_t ch = 0; ch < SC; ch++) {
                    const float wgt = weights[t * SC + ch];
                    if (wgt == 0.f) continue;
                    const float *tp = templates + (t * SC + ch) * L;
                    const float *d = data + ch * N;
                    const long mv = moveouts[t * SC + ch];
                    double ssd = 0.0;
                    for (size_t k = 0; k < L; k++) ssd += (double)d[a + mv + (long)k] * (double)d[a + mv + (long)k];
                    for (long n = a; n <= b; n++) {
                        const long j = n + mv;
                        const float num = dotf(tp, d + j, L);
                        const float den = sst[t * SC + ch] * (float)ssd;
                        float cc = 0.f;
                        if (den > MF_STABILITY_GUARD) cc = num / sqrtf(den);
                        o[n] += wgt * cc;
                        if (n < b) ssd += (double)d[j + (long)L] * (double)d[j + (long)L] - (double)d[j] * (double)d[j];
                    }
                }
            }
        }
    }
    free(sst);
    return 0;
}
