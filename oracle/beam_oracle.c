/*
 * oracle/beam_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (C99 + OpenMP) of the shift-and-stack network response that BPMF obtains from the
 * un-vendored pip dependency `beampower` (pyproject.toml:28 of the reference, no version pin,
 * absent from /root/reference and from this image).  PARITY UNPINNED at the kernel boundary; the
 * semantics follow SURVEY.md Appendix A.2, pinned by the reference's call sites and notebook math:
 *   - call / kwargs                BPMF/template_search.py:549-569
 *   - out_of_bounds meaning        BPMF/template_search.py:529-537
 *   - moveouts (K,S,P) int64 >= 0  BPMF/template_search.py:207-219 ; BPMF/utils.py:1270
 *   - weights_phases (S,C,P)       BPMF/template_search.py:766-769
 *   - weights_sources (K,S)        BPMF/template_search.py:770-772
 *   - (maxbeam, argmax) / (K,N)    BPMF/template_search.py:549 ; BPMF/dataset.py:2194-2196
 *   - U = sum_c alpha*F, b = sum_s beta * sum_p U(t+tau)   tutorial 5_backprojection.ipynb:876-882,928-929
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * call this file.
 *
 * Choices nothing pins: a (station, phase) term is "used" by source k iff
 * weights_sources[k,s] != 0; ties in argmax -> lowest source index; a sample where no source is
 * computable -> (0.0, 0).
 */
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <omp.h>

/* U[s,p,t] = sum_c wp[s,c,p] * F[s,c,t] */
void oracle_prestack(const float *features, const float *weights_phases, size_t S, size_t C,
                     size_t P, size_t N, float *U) {
#pragma omp parallel for collapse(2)
    for (size_t s = 0; s < S; s++)
        for (size_t p = 0; p < P; p++) {
            float *u = U + (s * P + p) * N;
            memset(u, 0, sizeof(float) * N);
            for (size_t c = 0; c < C; c++) {
                const float a = weights_phases[(s * C + c) * P + p];
                if (a == 0.f) continue;
                const float *f = features + (s * C + c) * N;
                for (size_t t = 0; t < N; t++) u[t] += a * f[t];
            }
        }
}

/* one beam sample; flexible != 0 drops out-of-range terms instead of voiding the sample */
static inline float beam_at(const float *U, const long long *tau, const float *beta, size_t S,
                            size_t P, size_t N, long long t, int flexible) {
    float b = 0.f;
    for (size_t s = 0; s < S; s++) {
        const float w = beta[s];
        if (w == 0.f) continue;
        for (size_t p = 0; p < P; p++) {
            const long long x = t + tau[s * P + p];
            if (x < 0 || x >= (long long)N) {
                if (flexible) continue;
                return 0.f;
            }
            b += w * U[(s * P + p) * N + (size_t)x];
        }
    }
    return b;
}

/*
 * reduce_max != 0: maxbeam f32[N], argmax i32[N];  beam unused (may be NULL).
 * reduce_max == 0: beam f32[K*N]; maxbeam/argmax unused.
 * k_offset is added to the reported argmax (source-sharded callers).
 */
int oracle_beamform(const float *features, const long long *moveouts, const float *weights_phases,
                    const float *weights_sources, size_t S, size_t C, size_t P, size_t N, size_t K,
                    int flexible, int reduce_max, int num_threads, float *maxbeam, int *argmax,
                    float *beam) {
    if (num_threads > 0) omp_set_num_threads(num_threads);
    float *U = (float *)malloc(sizeof(float) * S * P * N);
    if (!U) return -1;
    oracle_prestack(features, weights_phases, S, C, P, N, U);

    if (reduce_max) {
        /* time chunks in parallel, sources in ascending order inside a chunk so that the
         * strict ">" keeps the lowest source index on ties */
        const long long CH = 2048;
        const long long n_chunks = ((long long)N + CH - 1) / CH;
#pragma omp parallel for schedule(dynamic)
        for (long long q = 0; q < n_chunks; q++) {
            const long long t0 = q * CH;
            const long long t1 = (t0 + CH < (long long)N) ? t0 + CH : (long long)N;
            for (size_t k = 0; k < K; k++) {
                const long long *tau = moveouts + k * S * P;
                const float *beta = weights_sources + k * S;
                for (long long t = t0; t < t1; t++) {
                    float b = beam_at(U, tau, beta, S, P, N, t, flexible);
                    if (k == 0 || b > maxbeam[t]) { maxbeam[t] = b; argmax[t] = (int)k; }
                }
            }
        }
    } else {
#pragma omp parallel for schedule(static)
        for (long long k = 0; k < (long long)K; k++)
            for (long long t = 0; t < (long long)N; t++)
                beam[(size_t)k * N + (size_t)t] = beam_at(U, moveouts + (size_t)k * S * P,
                                                          weights_sources + (size_t)k * S, S, P, N,
                                                          t, flexible);
    }
    free(U);
    return 0;
}
