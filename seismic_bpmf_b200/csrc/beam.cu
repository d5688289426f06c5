// beam.cu -- shift-and-stack network response (beamforming / backprojection).
//
// Replaces the native kernels behind beampower.beampower.beamform as BPMF calls it
// (BPMF/template_search.py:549-569; math in tutorial 5_backprojection.ipynb:876-882,928-929);
// semantics: SURVEY.md A.2.
//
//   prestack : U[s,p,t] = sum_c weights_phases[s,c,p] * features[s,c,t]        (HBM-bound)
//   beam     : b[k,t]   = sum_{s,p} weights_sources[k,s] * U[s,p,t + tau[k,s,p]]
//              CTA = 2048-sample time tile x a chunk of sources, processed 8 sources at a time
//              with 8x8 register accumulators per thread.  The (station, phase) loop is OUTER: one
//              row segment of U per (station, phase) -- one window covers the block's 8 sources when
//              their moveouts are close, as for neighbouring grid points -- streams through a
//              cp.async double buffer while the previous row is consumed; every gather-FMA costs
//              one conflict-free 4-byte LDS (lanes = consecutive samples).  Roofline: shared-memory
//              bandwidth, 128 B/clk/SM = 32 gather-FMA/clk/SM.
//   reduce   : "max" keeps a running (max, argmax) per sample in registers and merges source
//              chunks with a packed 64-bit atomicMax (value-ordered bits << 32 | ~k, so ties go to
//              the lowest source index); "none" stores b[k,t].
#include <limits.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "async.cuh"
#include "common.cuh"

namespace bpmf {

constexpr int B_THREADS = 256;
constexpr int B_R = 8;                    // samples per thread
constexpr int B_TT = B_THREADS * B_R;     // 2048-sample tile
constexpr int B_KB = 8;                   // sources per register block
constexpr int B_CAP = B_TT + 1024;        // staged row segment (floats)
constexpr int B_NBUF = 3;                 // row segments in flight

// U rows have a pitch that is a multiple of 4 floats (16-byte aligned rows for the float4 staging of the
// beam kernel); the pad beyond N is zero.
__global__ void beam_prestack_kernel(const float *__restrict__ features,
                                     const float *__restrict__ wp, int C, int P, size_t N, size_t pitch,
                                     float *__restrict__ U) {
    const int sp = blockIdx.y;  // s * P + p
    const int s = sp / P, p = sp % P;
    const size_t t = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (t >= pitch) return;
    float a[4] = {0.f, 0.f, 0.f, 0.f};
    const int n_here = (t >= N) ? 0 : ((N - t >= 4) ? 4 : (int)(N - t));
    for (int c = 0; c < C; c++) {
        const float w = wp[((size_t)s * C + c) * P + p];
        if (w == 0.f) continue;
        const float *f = features + ((size_t)s * C + c) * N + t;
        for (int i = 0; i < n_here; i++) a[i] = fmaf(w, f[i], a[i]);
    }
    float *u = U + (size_t)sp * pitch + t;
    for (int i = 0; i < 4 && t + i < pitch; i++) u[i] = a[i];
}

__device__ __forceinline__ unsigned ordered_bits(float f) {
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float from_ordered_bits(unsigned u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// per-source valid time range for out_of_bounds="strict": [tlo, thi] inclusive
__global__ void beam_source_bounds_kernel(const int *__restrict__ tau, const float *__restrict__ beta,
                                          size_t K, int S, int P, long long N, int *__restrict__ tlo,
                                          int *__restrict__ thi) {
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    long long mn = 0, mx = 0;
    bool any = false;
    for (int s = 0; s < S; s++) {
        if (beta[k * S + s] == 0.f) continue;
        for (int p = 0; p < P; p++) {
            const long long m = tau[(k * S + s) * P + p];
            if (!any) { mn = mx = m; any = true; }
            mn = m < mn ? m : mn;
            mx = m > mx ? m : mx;
        }
    }
    long long lo = 0, hi = N - 1;
    if (any) {
        if (mn < 0) lo = -mn;
        hi = N - 1 - mx;
        if (hi > N - 1) hi = N - 1;
    }
    tlo[k] = (int)lo;
    thi[k] = (int)(hi < -1 ? -1 : hi);
}

// 16-byte async copy global -> shared with zero fill when `valid` is false (LDGSTS.128)
__device__ __forceinline__ void cp_async_f32x4(float *smem_dst, const float *gmem_src, bool valid) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int src_bytes = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(gmem_src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// stage samples [first, first + B_CAP) of one U row into `buf`, zero outside the row; `first` and the row
// pitch are multiples of 4, so every 16-byte piece is entirely inside or entirely outside [0, pitch)
__device__ __forceinline__ void beam_stage_async(float *buf, const float *row, long long first, long long pitch,
                                                 int tid) {
#pragma unroll
    for (int i = tid * 4; i < B_CAP; i += B_THREADS * 4) {
        const long long g = first + i;
        const bool ok = g >= 0 && g < pitch;
        cp_async_f32x4(buf + i, row + (ok ? g : 0), ok);
    }
}

template <bool REDUCE_MAX>
__global__ void __launch_bounds__(B_THREADS, 2)
    beam_kernel(const float *__restrict__ U, const int *__restrict__ tau,
                const float *__restrict__ beta, const int *__restrict__ tlo,
                const int *__restrict__ thi, size_t K, int S, int P, long long N, long long pitch, int strict,
                size_t k_chunk, unsigned long long *__restrict__ keys, float *__restrict__ beam) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *buf0 = reinterpret_cast<float *>(smem_raw);                 // B_NBUF x B_CAP (ring of row segments)
    int *tau_s = reinterpret_cast<int *>(buf0 + B_NBUF * B_CAP);       // B_KB * S * P
    float *beta_s = reinterpret_cast<float *>(tau_s + B_KB * S * P);   // B_KB * S
    int *wlo_s = reinterpret_cast<int *>(beta_s + B_KB * S);           // S * P: window start per (station, phase)
    int *wide_s = wlo_s + S * P;                                       // S * P: 1 = moveouts do not fit one window
    // S * P * B_KB packed (offset in window, weight bits), 16-byte aligned for the broadcast LDS.128
    int2 *pk_s = reinterpret_cast<int2 *>((reinterpret_cast<uintptr_t>(wide_s + S * P) + 15) & ~(uintptr_t)15);

    const int tid = threadIdx.x;
    const long long t0 = (long long)blockIdx.x * B_TT;
    const size_t k_begin = (size_t)blockIdx.y * k_chunk;
    const size_t k_end = (k_begin + k_chunk < K) ? k_begin + k_chunk : K;
    const int SP = S * P;

    float best[B_R];
    unsigned bestk[B_R];
#pragma unroll
    for (int r = 0; r < B_R; r++) { best[r] = -INFINITY; bestk[r] = 0u; }

    for (size_t kb = k_begin; kb < k_end; kb += B_KB) {
        const int nk = (int)((k_end - kb < (size_t)B_KB) ? (k_end - kb) : (size_t)B_KB);
        __syncthreads();
        for (int i = tid; i < B_KB * SP; i += B_THREADS) {
            const int kk = i / SP;
            tau_s[i] = (kk < nk) ? tau[(kb + kk) * SP + (i - kk * SP)] : 0;
        }
        for (int i = tid; i < B_KB * S; i += B_THREADS) {
            const int kk = i / S;
            beta_s[i] = (kk < nk) ? beta[(kb + kk) * S + (i - kk * S)] : 0.f;
        }
        __syncthreads();
        // one staged window per (station, phase) covers every source of the block when their moveouts
        // are close (neighbouring grid points); otherwise the slow path restages per source
        for (int sp = tid; sp < SP; sp += B_THREADS) {
            const int s = sp / P;
            int mn = 0x7fffffff, mx = -0x7fffffff;
            for (int kk = 0; kk < B_KB; kk++) {
                if (beta_s[kk * S + s] == 0.f) continue;
                const int v = tau_s[kk * SP + sp];
                mn = v < mn ? v : mn;
                mx = v > mx ? v : mx;
            }
            if (mx < mn) { mn = 0; mx = 0; wide_s[sp] = 2; }  // no source uses this row
            else {
                mn = (mn >= 0) ? (mn & ~3) : -((-mn + 3) & ~3);  // align down: 16-byte staging
                wide_s[sp] = ((long long)mx - mn + B_TT > B_CAP) ? 1 : 0;
            }
            wlo_s[sp] = mn;
            for (int kk = 0; kk < B_KB; kk++)
                pk_s[sp * B_KB + kk] = make_int2(tau_s[kk * SP + sp] - mn, __float_as_int(beta_s[kk * S + s]));
        }
        __syncthreads();

        float acc[B_KB][B_R];
#pragma unroll
        for (int kk = 0; kk < B_KB; kk++)
#pragma unroll
            for (int r = 0; r < B_R; r++) acc[kk][r] = 0.f;

        // software pipeline over (station, phase): rows sp+1 and sp+2 stream in with cp.async while row sp is
        // used (3-slot ring, ONE barrier per row: it publishes row sp and retires row sp-1, whose slot is then
        // refilled with row sp+2)
#pragma unroll
        for (int pre = 0; pre < 2; pre++) {
            if (pre < SP && wide_s[pre] != 2)
                beam_stage_async(buf0 + pre * B_CAP, U + (size_t)pre * pitch, t0 + wlo_s[pre], pitch, tid);
            cp_async_commit();
        }
        for (int sp = 0; sp < SP; sp++) {
            const int s = sp / P;
            float *buf = buf0 + (sp % B_NBUF) * B_CAP;
            cp_async_wait<1>();  // row sp has landed (row sp+1 may still be in flight)
            __syncthreads();
            if (sp + 2 < SP && wide_s[sp + 2] != 2)
                beam_stage_async(buf0 + ((sp + 2) % B_NBUF) * B_CAP, U + (size_t)(sp + 2) * pitch, t0 + wlo_s[sp + 2],
                                 pitch, tid);
            cp_async_commit();
            const int wide = wide_s[sp];
            if (wide == 0) {
                // straight-line code over the block's 8 sources: (offset, weight) pairs come as broadcast
                // LDS.128, zero-weight (source, station) pairs are predicated off instead of branched over
                int2 pk[B_KB];
                const int4 *pk4 = reinterpret_cast<const int4 *>(pk_s + sp * B_KB);
#pragma unroll
                for (int q = 0; q < B_KB / 2; q++) {
                    const int4 v = pk4[q];
                    pk[2 * q] = make_int2(v.x, v.y);
                    pk[2 * q + 1] = make_int2(v.z, v.w);
                }
#pragma unroll
                for (int kk = 0; kk < B_KB; kk++) {
                    const float b = __int_as_float(pk[kk].y);
                    const float *src = buf + pk[kk].x + tid;
#pragma unroll
                    for (int r = 0; r < B_R; r++) {
                        const float u = src[B_THREADS * r];
                        acc[kk][r] = (b != 0.f) ? fmaf(b, u, acc[kk][r]) : acc[kk][r];
                    }
                }
            } else if (wide == 1) {
                // moveouts of this block spread beyond one window: restage per source (synchronous)
                const float *row = U + (size_t)sp * pitch;
                long long wlo = wlo_s[sp];
#pragma unroll
                for (int kk = 0; kk < B_KB; kk++) {
                    const float b = beta_s[kk * S + s];
                    if (b == 0.f) continue;
                    const long long tk = tau_s[kk * SP + sp];
                    if (tk < wlo || tk + B_TT > wlo + B_CAP) {
                        __syncthreads();
                        wlo = tk;
                        for (int i = tid; i < B_CAP; i += B_THREADS) {
                            const long long g = t0 + wlo + i;
                            buf[i] = (g >= 0 && g < N) ? __ldg(row + g) : 0.f;
                        }
                        __syncthreads();
                    }
                    const float *src = buf + (int)(tk - wlo) + tid;
#pragma unroll
                    for (int r = 0; r < B_R; r++) acc[kk][r] = fmaf(b, src[B_THREADS * r], acc[kk][r]);
                }
            }
        }
        cp_async_wait<0>();

#pragma unroll
        for (int kk = 0; kk < B_KB; kk++) {
            if (kk < nk) {
                const size_t k = kb + kk;
                long long lo_k = 0, hi_k = N - 1;
                if (strict) { lo_k = tlo[k]; hi_k = thi[k]; }
#pragma unroll
                for (int r = 0; r < B_R; r++) {
                    const long long t = t0 + tid + B_THREADS * r;
                    const float v = (t >= lo_k && t <= hi_k) ? acc[kk][r] : 0.f;
                    if (REDUCE_MAX) {
                        if (v > best[r]) { best[r] = v; bestk[r] = (unsigned)k; }
                    } else if (t < N) {
                        beam[k * (size_t)N + (size_t)t] = v;
                    }
                }
            }
        }
    }
    if (REDUCE_MAX && k_end > k_begin) {
#pragma unroll
        for (int r = 0; r < B_R; r++) {
            const long long t = t0 + tid + B_THREADS * r;
            if (t < N) {
                const unsigned long long key =
                    ((unsigned long long)ordered_bits(best[r]) << 32) | (unsigned long long)(~bestk[r]);
                atomicMax(keys + t, key);
            }
        }
    }
}

// ---------------------------------------------------------------- grouped, warp-specialised beam kernel
// The sources are cut into super-blocks of 16 consecutive sources; inside a super-block, consecutive
// sources are grouped greedily as long as, for EVERY (station, phase) row, the moveouts of the group's
// weighted members fit one staged window (spread <= F_MARGIN samples).  Neighbouring grid points
// give groups of 16; unrelated moveouts degrade to groups of one source, never to a different code path.
//   beam_pack_kernel : warp per super-block -> transposed moveouts tauT[sb][sp][16], weights
//                      betaT[sb][s][16] and the group list grp[sb][16] (start | count << 8, 0 = end).
//   beam_group_kernel: CTA = 2688-sample time tile x a range of super-blocks, ONE CTA per SM.  One PRODUCER
//                      warp walks (group, row): window start = min moveout of the group, (offset, weight)
//                      table of the row, one cp.async.bulk (TMA) of the <= 12.5 KB row window into a ring of
//                      F_NSTAGE stages, completion on the stage's "full" mbarrier.  Twelve CONSUMER warps
//                      (8 sources x 14 samples of register accumulators per thread, lanes = consecutive
//                      samples: conflict-free 4-byte LDS per gather-FMA) wait on "full", accumulate,
//                      and release the stage on its "empty" mbarrier -- no CTA-wide barrier, warps drift.
// Tile shape (r2 sweeps, gpurun r2s25-r2s35; bench grid, dense / 10-closest weighting, ms per pass):
//   3 samples x 8 warps x 2 CTAs/SM (r1)        269 / 126      6 x 10 warps, 1 CTA/SM              208 / 92
//   10 x 8 warps, setmaxnreg 232                196 / 73.5     + 2 warps share the samples (SPLIT 2) 191 / 73.5
//   + unused rows skipped 32 at a time, valid ranges in the terminator, running max in shared memory,
//     7 x 12 warps, setmaxnreg 152 (this file)  189 / 60
// The loop is bound by the gathers in flight: ptxas keeps a sliding window of ~11 loads ahead of the first FMA
// and issue is in order, so the warps (12 x 11 loads) have to cover the shared-memory latency under load; the
// samples per thread amortise the row hand-shake (wait -> header -> table -> first gather, ~120 cycles) and the
// 8 table wavefronts of the row.  Both are bought with registers, hence one CTA per SM and setmaxnreg.
#ifndef BEAM_CONS_WARPS
#define BEAM_CONS_WARPS 12
#endif
constexpr int F_CONS_WARPS = BEAM_CONS_WARPS;
constexpr int F_CONS = 32 * F_CONS_WARPS;        // consumer threads
// BEAM_CONS_REGS != 0: the producer sits in a warpgroup of its own (its three other warps leave at once) and hands
// its registers to the three consumer warpgroups (setmaxnreg): 128 registers per thread at launch (512 threads, one
// CTA per SM), then 40 for the producer warpgroup and 152 for the consumers -- 112 accumulators + gathers in flight.
#ifndef BEAM_CONS_REGS
#define BEAM_CONS_REGS 152
#endif
#ifndef BEAM_PROD_REGS
#define BEAM_PROD_REGS 40
#endif
constexpr int F_THREADS = F_CONS + (BEAM_CONS_REGS ? 128 : 32);  // + the producer warp (warpgroup)
static_assert(!BEAM_CONS_REGS || F_CONS_WARPS % 4 == 0, "setmaxnreg works on whole warpgroups");
// SPLIT (template parameter of the kernel) = consumer warps that share the samples of a group, each taking
// 16 / SPLIT of its sources and BEAM_SPT * SPLIT samples per thread: always 16 * BEAM_SPT accumulators.  SPLIT = 2
// (half the table loads per gather) is the default; BPMF_BEAM_SPLIT=1 selects the other instantiation for A/B runs.
#ifndef BEAM_SPT
#define BEAM_SPT 7                               // samples per consumer thread (x SPLIT): 112 accumulators
#endif
#ifndef BEAM_MINB
#define BEAM_MINB 1                              // CTAs per SM the register budget is sized for
#endif
constexpr int F_TT = F_CONS * BEAM_SPT;          // 2688-sample tile
constexpr int F_G = 16;                          // sources per super-block / largest group
#ifndef BEAM_DENSE_MIN
#define BEAM_DENSE_MIN 8                         // rows with more weighted positions than this run the straight-line code
#endif
#ifndef BEAM_MARGIN
#define BEAM_MARGIN 512
#endif
#ifndef BEAM_NSTAGE
#define BEAM_NSTAGE 8
#endif
constexpr int F_MARGIN = BEAM_MARGIN;                    // largest moveout spread inside a group (samples)
constexpr int F_WIN = F_TT + F_MARGIN;           // staged row window (floats)
constexpr int F_NSTAGE = BEAM_NSTAGE;
constexpr int F_DIRECT_MAX = 2;                  // groups of up to two sources are gathered straight from global memory
constexpr int F_STAGE = F_WIN + 2 * F_G + 8;     // + (offset, weight) table + header (skip flag), floats; x4 B is a multiple of 16

__device__ __forceinline__ long long align4_down(long long v) { return v & ~3ll; }

__global__ void __launch_bounds__(128)
    beam_pack_kernel(const int *__restrict__ tau, const float *__restrict__ beta, size_t K, int S, int P,
                     int *__restrict__ tauT, float *__restrict__ betaT, unsigned short *__restrict__ grp,
                     int *__restrict__ dlist, int *__restrict__ dcount) {
    // dcount[0]: sources listed for the direct kernel
    extern __shared__ long long pack_sm[];  // per warp: mn[SP], mx[SP]
    const int SP = S * P;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const size_t sb = (size_t)blockIdx.x * (blockDim.x >> 5) + wib;
    const size_t n_sb = (K + F_G - 1) / F_G;
    if (sb >= n_sb) return;
    long long *mn_s = pack_sm + (size_t)wib * 2 * SP, *mx_s = mn_s + SP;
    const size_t k0 = sb * F_G;
    const int nk = (int)((K - k0 < (size_t)F_G) ? (K - k0) : (size_t)F_G);
    // transposes (reads are contiguous: 16 consecutive sources)
    for (int e = lane; e < F_G * SP; e += 32) {
        const int j = e / SP, sp = e - j * SP;
        tauT[(sb * SP + sp) * F_G + j] = (j < nk) ? tau[(k0 + j) * SP + sp] : 0;
    }
    for (int e = lane; e < F_G * S; e += 32) {
        const int j = e / S, s = e - j * S;
        betaT[(sb * S + s) * F_G + j] = (j < nk) ? beta[(k0 + j) * S + s] : 0.f;
    }
    // greedy grouping
    for (int sp = lane; sp < SP; sp += 32) { mn_s[sp] = LLONG_MAX; mx_s[sp] = LLONG_MIN; }
    int g_count = 0, g_start = 0, g_n = 0;
    for (int j = 0; j < nk; j++) {
        bool bad = false;
        for (int sp = lane; sp < SP; sp += 32) {
            if (beta[(k0 + j) * S + sp / P] == 0.f) continue;
            const long long v = tau[(k0 + j) * SP + sp];
            const long long a = v < mn_s[sp] ? v : mn_s[sp], b = v > mx_s[sp] ? v : mx_s[sp];
            if (b - align4_down(a) > F_MARGIN) bad = true;
        }
        if (__any_sync(0xffffffffu, bad)) {  // close the group, source j starts the next one
            if (lane == 0) grp[sb * F_G + g_count] = (unsigned short)(g_start | (g_n << 8));
            g_count++;
            g_start = j;
            g_n = 0;
            for (int sp = lane; sp < SP; sp += 32) { mn_s[sp] = LLONG_MAX; mx_s[sp] = LLONG_MIN; }
        }
        for (int sp = lane; sp < SP; sp += 32) {
            if (beta[(k0 + j) * S + sp / P] == 0.f) continue;
            const long long v = tau[(k0 + j) * SP + sp];
            if (v < mn_s[sp]) mn_s[sp] = v;
            if (v > mx_s[sp]) mx_s[sp] = v;
        }
        g_n++;
    }
    if (lane == 0) {
        if (g_n > 0) grp[sb * F_G + g_count++] = (unsigned short)(g_start | (g_n << 8));
        for (int g = g_count; g < F_G; g++) grp[sb * F_G + g] = 0;
        // sources of the groups too small for a staged window go to beam_direct_kernel
        for (int g = 0; g < g_count; g++) {
            const unsigned gd = grp[sb * F_G + g];
            const int n = (int)(gd >> 8), j0 = (int)(gd & 0xffu);
            if (n > F_DIRECT_MAX) continue;
            const int pos = atomicAdd(dcount, n);
            for (int i = 0; i < n; i++) dlist[pos + i] = (int)(k0 + j0 + i);
        }
    }
}

// Record stream producer -> consumers (one record per stage): a record is one (station, phase) row that at
// least one member of the group uses -- rows nobody uses are never emitted -- and every group ends with a
// terminator record.  Record kinds: F_ROW_DENSE (straight-line code over the 16 positions; positions without
// weight point at a block of zeros with weight 0, so that fmaf(0, 0, acc) leaves acc untouched whatever the
// row holds), F_ROW_SPARSE (at most half of the positions weighted: warp-uniform branches), F_END / F_END_INSIDE.
// The terminator carries, per source of the group, the range of tile samples that are valid for it (strict mode:
// every station's window inside the data; read from tlo / thi by the producer, which runs ahead, instead of
// sixteen dependent global loads in every consumer at the end of every group); F_END_INSIDE = every sample of the
// tile is valid for every source, the consumers skip the range tests.
// Groups of one or two sources (unrelated moveouts) are not worth a staged window: beam_pack_kernel lists
// their sources and beam_direct_kernel gathers them straight from the L2-resident rows (4 B per gather-FMA
// either way, without the per-row hand-shake); both kernels merge into the same (max, argmax) keys.
#ifdef BEAM_DIAG
// diagnostic build only (tools/gpu scripts): cycles spent waiting, summed over CTAs
//   [0] producer: waiting for a free stage   [1] producer: waiting for super-block parameters   [2] producer: total
//   [3] consumer warp 0: waiting for a full stage   [4] consumer warp 0: group epilogues   [5] consumer warp 0: total
//   [6] records   [7] groups
__device__ unsigned long long beam_diag[8];
#define DIAG_T0(v) const long long v = clock64()
#define DIAG_ADD(acc, v) acc += clock64() - v
#else
#define DIAG_T0(v)
#define DIAG_ADD(acc, v)
#endif
constexpr int F_ROW_DENSE = 0, F_ROW_SPARSE = 2, F_END = 1, F_END_INSIDE = 3;  // bit 0: terminator

template <bool REDUCE_MAX, int SPLIT>
__global__ void __launch_bounds__(F_THREADS, BEAM_MINB)
    beam_group_kernel(const float *__restrict__ U, const int *__restrict__ tauT, const float *__restrict__ betaT,
                      const unsigned short *__restrict__ grp, const int *__restrict__ tlo,
                      const int *__restrict__ thi, size_t K, int S, int P, long long N, long long pitch, int strict,
                      size_t sb_chunk, unsigned long long *__restrict__ keys, float *__restrict__ beam) {
    constexpr int F_R = BEAM_SPT * SPLIT, F_LANES = F_CONS / SPLIT, F_GW = F_G / SPLIT;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int SP = S * P;
    float *stages = reinterpret_cast<float *>(smem_raw);                        // F_NSTAGE x F_STAGE
    float *zero_s = stages + F_NSTAGE * F_STAGE;                                // F_TT zeros
    int *par_tau = reinterpret_cast<int *>(zero_s + F_TT);                      // 2 x SP x 16
    float *par_beta = reinterpret_cast<float *>(par_tau + 2 * SP * F_G);        // 2 x S x 16
    uint64_t *bars = reinterpret_cast<uint64_t *>(par_beta + 2 * S * F_G);
    uint64_t *full = bars, *empty = bars + F_NSTAGE, *pfull = bars + 2 * F_NSTAGE;  // [F_NSTAGE], [F_NSTAGE], [2]
    float *best_s = reinterpret_cast<float *>(bars + 2 * F_NSTAGE + 2);                // REDUCE_MAX: [F_R][F_CONS]
    unsigned *bestk_s = reinterpret_cast<unsigned *>(best_s + F_R * F_CONS);           //             [F_R][F_CONS]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long t0 = (long long)blockIdx.x * F_TT;
    const size_t n_sb = (K + F_G - 1) / F_G;
    const size_t sb_begin = (size_t)blockIdx.y * sb_chunk;
    const size_t sb_end = (sb_begin + sb_chunk < n_sb) ? sb_begin + sb_chunk : n_sb;

    for (int i = tid; i < F_TT; i += F_THREADS) zero_s[i] = 0.f;
    if (tid == 0) {
        for (int i = 0; i < F_NSTAGE; i++) {
            mbar_init(full + i, 1);
            mbar_init(empty + i, F_CONS_WARPS);
        }
        mbar_init(pfull + 0, 1);
        mbar_init(pfull + 1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
#if BEAM_CONS_REGS
    if (warp >= F_CONS_WARPS) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(BEAM_PROD_REGS));
        if (warp != F_CONS_WARPS) return;
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(BEAM_CONS_REGS));
    }
#endif

    if (warp == F_CONS_WARPS) {
        // ===================== producer warp ==========================================================
        const uint32_t par_bytes = (uint32_t)(sizeof(int) * SP * F_G + sizeof(float) * S * F_G);
        uint32_t it = 0;
#ifdef BEAM_DIAG
        long long d_par = 0, d_empty = 0, d_groups = 0;
        const long long d_start = clock64();
#endif
        for (size_t sb = sb_begin; sb < sb_end; sb++) {
            const uint32_t u = (uint32_t)(sb - sb_begin);
            const int pb = u & 1;
            __syncwarp();
            if (lane == 0) {
                if (u == 0) {
                    mbar_expect_tx(pfull + pb, par_bytes);
                    bulk_g2s(par_tau + pb * SP * F_G, tauT + sb * SP * F_G, (uint32_t)(sizeof(int) * SP * F_G), pfull + pb);
                    bulk_g2s(par_beta + pb * S * F_G, betaT + sb * S * F_G, (uint32_t)(sizeof(float) * S * F_G), pfull + pb);
                }
                if (sb + 1 < sb_end) {  // the other buffer was last read for super-block sb - 1
                    mbar_expect_tx(pfull + (pb ^ 1), par_bytes);
                    bulk_g2s(par_tau + (pb ^ 1) * SP * F_G, tauT + (sb + 1) * SP * F_G,
                             (uint32_t)(sizeof(int) * SP * F_G), pfull + (pb ^ 1));
                    bulk_g2s(par_beta + (pb ^ 1) * S * F_G, betaT + (sb + 1) * S * F_G,
                             (uint32_t)(sizeof(float) * S * F_G), pfull + (pb ^ 1));
                }
            }
            {
                DIAG_T0(c0);
                mbar_wait(pfull + pb, (u >> 1) & 1);
                DIAG_ADD(d_par, c0);
            }
            const int *tau_p = par_tau + pb * SP * F_G;
            const float *beta_p = par_beta + pb * S * F_G;
            for (int g = 0; g < F_G; g++) {
                const unsigned gd = grp[sb * F_G + g];
                const int n = (int)(gd >> 8), j0 = (int)(gd & 0xffu);
                if (n == 0) break;
                if (n <= F_DIRECT_MAX) continue;  // left to beam_direct_kernel
                // valid samples of source j0 + lane, relative to the tile (used by the terminator record)
                long long lo_k = 0, hi_k = N - 1;
                if (strict && lane < n) {
                    lo_k = tlo[sb * F_G + j0 + lane];
                    hi_k = thi[sb * F_G + j0 + lane];
                }
                for (int sp0 = 0; sp0 < SP; sp0 += 32) {
                  // rows sp0 .. sp0 + 31 that at least one member of the group uses (one lane per row); the others
                  // are never emitted, and skipping them 32 at a time keeps the producer ahead of the consumers
                  // when every source weights only a few stations
                  bool used = false;
                  if (sp0 + lane < SP) {
                      const float *bp = beta_p + ((sp0 + lane) / P) * F_G + j0;
                      for (int a = 0; a < n; a++) used |= bp[a] != 0.f;
                  }
                  unsigned row_mask = __ballot_sync(0xffffffffu, used);
                  while (row_mask) {
                    const int sp = sp0 + __ffs(row_mask) - 1;
                    row_mask &= row_mask - 1;
                    const int s = sp / P;
                    // lane a < 16 owns position a of the group (source slot j0 + a)
                    float b = 0.f;
                    int tv = 0;
                    if (lane < n) {
                        b = beta_p[s * F_G + j0 + lane];
                        tv = tau_p[sp * F_G + j0 + lane];
                    }
                    const unsigned nz = __ballot_sync(0xffffffffu, b != 0.f);
                    int mn = (b != 0.f) ? tv : INT_MAX, mx = (b != 0.f) ? tv : INT_MIN;
#pragma unroll
                    for (int o = 8; o > 0; o >>= 1) {  // lanes 16-31 hold the neutral elements
                        const int x = __shfl_xor_sync(0xffffffffu, mn, o), y = __shfl_xor_sync(0xffffffffu, mx, o);
                        mn = x < mn ? x : mn;
                        mx = y > mx ? y : mx;
                    }
                    mn = __shfl_sync(0xffffffffu, mn, 0);
                    mx = __shfl_sync(0xffffffffu, mx, 0);
                    const int stage = it % F_NSTAGE;
                    {
                        DIAG_T0(c0);
                        mbar_wait(empty + stage, ((it / F_NSTAGE) & 1) ^ 1);
                        DIAG_ADD(d_empty, c0);
                    }
                    it++;
                    float *st = stages + stage * F_STAGE;
                    int *hdr = reinterpret_cast<int *>(st + F_WIN + 2 * F_G);
                    const long long wlo = align4_down((long long)mn);
                    const int dense = __popc(nz) > BEAM_DENSE_MIN;
                    if (lane < F_G)
                        reinterpret_cast<int2 *>(st + F_WIN)[lane] =
                            (b != 0.f) ? make_int2(4 * (int)((long long)tv - wlo), __float_as_int(b))
                                       : make_int2(4 * (int)(zero_s - st), 0);  // BYTE offsets: one add per gather address
                    if (lane == 0) hdr[0] = dense ? F_ROW_DENSE : F_ROW_SPARSE;
                    // samples [g0, g0 + len) of the row are needed; the part outside [0, pitch) reads as zero
                    const long long g0 = t0 + wlo;                                      // multiple of 4
                    const int len = (int)(((long long)mx - wlo + F_TT + 3) & ~3ll);     // <= F_WIN
                    long long a0 = g0 < 0 ? 0 : g0;
                    long long b0 = (g0 + len < pitch) ? g0 + len : pitch;
                    if (b0 < a0) a0 = b0 = g0 + len;
                    const int za = (int)(a0 - g0), zb = (int)(b0 - g0);  // in-range part of the window: [za, zb)
                    if (za > 0 || zb < len) {                            // edges of the record only
                        for (int i = 4 * lane; i < len; i += 128)
                            if (i < za || i >= zb) *reinterpret_cast<float4 *>(st + i) = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                    __syncwarp();
                    if (lane == 0) {
                        if (zb > za) {
                            mbar_expect_tx(full + stage, (uint32_t)(zb - za) * 4u);
                            bulk_g2s(st + za, U + (size_t)sp * pitch + a0, (uint32_t)(zb - za) * 4u, full + stage);
                        } else {
                            mbar_arrive(full + stage);
                        }
                    }
                  }
                }
                // terminator record of the group
                const int stage = it % F_NSTAGE;
                mbar_wait(empty + stage, ((it / F_NSTAGE) & 1) ^ 1);
                it++;
                {
                    float *st = stages + stage * F_STAGE;
                    int lo_rel = F_TT, hi_rel = -1;  // positions beyond the group: empty range
                    if (lane < n) {
                        const long long a = lo_k - t0, b = hi_k - t0;
                        lo_rel = (int)(a < 0 ? 0 : (a > F_TT ? F_TT : a));
                        hi_rel = (int)(b < -1 ? -1 : (b > F_TT - 1 ? F_TT - 1 : b));
                    }
                    const bool inside = __all_sync(0xffffffffu, lane >= n || (lo_rel == 0 && hi_rel == F_TT - 1));
                    if (lane < F_G) reinterpret_cast<int2 *>(st + F_WIN)[lane] = make_int2(lo_rel, hi_rel);
                    if (lane == 0) reinterpret_cast<int *>(st + F_WIN + 2 * F_G)[0] = inside ? F_END_INSIDE : F_END;
                    __syncwarp();
                    if (lane == 0) mbar_arrive(full + stage);
                }
#ifdef BEAM_DIAG
                d_groups++;
#endif
            }
        }
#ifdef BEAM_DIAG
        if (lane == 0) {
            atomicAdd(beam_diag + 0, (unsigned long long)d_empty);
            atomicAdd(beam_diag + 1, (unsigned long long)d_par);
            atomicAdd(beam_diag + 2, (unsigned long long)(clock64() - d_start));
            atomicAdd(beam_diag + 6, (unsigned long long)it);
            atomicAdd(beam_diag + 7, (unsigned long long)d_groups);
        }
#endif
    } else {
        // ===================== consumer warps =========================================================
        // warp w takes source positions [F_GW * (w % SPLIT), + F_GW) of every group and the samples
        // t0 + ts + F_LANES * r, ts = 32 * (w / SPLIT) + lane, r < F_R (lanes = consecutive samples)
        const int half = warp % SPLIT;
        const int ts = 32 * (warp / SPLIT) + lane;
        if (REDUCE_MAX) {
#pragma unroll
            for (int r = 0; r < F_R; r++) { best_s[r * F_CONS + tid] = -INFINITY; bestk_s[r * F_CONS + tid] = 0u; }
        }
        uint32_t it = 0;
        bool any = false;
#ifdef BEAM_DIAG
        long long d_full = 0, d_epi = 0;
        const long long d_start = clock64();
#endif
        for (size_t sb = sb_begin; sb < sb_end; sb++) {
            for (int g = 0; g < F_G; g++) {
                const unsigned gd = grp[sb * F_G + g];
                const int n = (int)(gd >> 8), j0 = (int)(gd & 0xffu);
                if (n == 0) break;
                if (n <= F_DIRECT_MAX) continue;  // beam_direct_kernel gathers these straight from global memory
                any = true;
                float acc[F_GW][F_R];
#pragma unroll
                for (int a = 0; a < F_GW; a++)
#pragma unroll
                    for (int r = 0; r < F_R; r++) acc[a][r] = 0.f;
                int stage, kind;
                const float *st;
                for (;;) {
                    stage = it % F_NSTAGE;
                    {
                        DIAG_T0(c0);
                        mbar_wait(full + stage, (it / F_NSTAGE) & 1);
                        DIAG_ADD(d_full, c0);
                    }
                    it++;
                    st = stages + stage * F_STAGE;
                    kind = reinterpret_cast<const int *>(st + F_WIN + 2 * F_G)[0];
                    // (offset, weight) of this warp's positions 2q, 2q + 1
                    const int4 *pk4 = reinterpret_cast<const int4 *>(st + F_WIN) + half * (F_GW / 2);
                    const char *sb = reinterpret_cast<const char *>(st + ts);  // this thread's first sample of the window
                    if (kind == F_ROW_DENSE) {
#pragma unroll
                        for (int q = 0; q < F_GW / 2; q++) {
                            const int4 v = pk4[q];
                            const float b0 = __int_as_float(v.y), b1 = __int_as_float(v.w);
                            const float *s0 = reinterpret_cast<const float *>(sb + v.x);
                            const float *s1 = reinterpret_cast<const float *>(sb + v.z);
#pragma unroll
                            for (int r = 0; r < F_R; r++) acc[2 * q][r] = fmaf(b0, s0[F_LANES * r], acc[2 * q][r]);
#pragma unroll
                            for (int r = 0; r < F_R; r++) acc[2 * q + 1][r] = fmaf(b1, s1[F_LANES * r], acc[2 * q + 1][r]);
                        }
                    } else if (kind == F_ROW_SPARSE) {
#pragma unroll
                        for (int q = 0; q < F_GW / 2; q++) {
                            const int4 v = pk4[q];
                            const float b0 = __int_as_float(v.y), b1 = __int_as_float(v.w);
                            if (b0 != 0.f) {
                                const float *s0 = reinterpret_cast<const float *>(sb + v.x);
#pragma unroll
                                for (int r = 0; r < F_R; r++) acc[2 * q][r] = fmaf(b0, s0[F_LANES * r], acc[2 * q][r]);
                            }
                            if (b1 != 0.f) {
                                const float *s1 = reinterpret_cast<const float *>(sb + v.z);
#pragma unroll
                                for (int r = 0; r < F_R; r++) acc[2 * q + 1][r] = fmaf(b1, s1[F_LANES * r], acc[2 * q + 1][r]);
                            }
                        }
                    }
                    if (kind & 1) break;  // terminator: released after the ranges have been read
                    __syncwarp();
                    if (lane == 0) mbar_arrive(empty + stage);
                }
                DIAG_T0(c_epi);
                const int2 *rng = reinterpret_cast<const int2 *>(st + F_WIN) + half * F_GW;
                const int na = n - half * F_GW;  // this warp's positions [0, na) hold sources of the group
                if (REDUCE_MAX) {
                    // running (max, argmax) of the thread's samples: in shared memory, read and written once per
                    // group -- 2 * F_R registers less in the gather loop, i.e. more gathers in flight
                    const unsigned kbase = (unsigned)(sb * F_G + j0 + half * F_GW);
                    if (kind == F_END_INSIDE) {  // every sample of the tile is valid for every source of the group
#pragma unroll
                        for (int r = 0; r < F_R; r++) {
                            float bv = best_s[r * F_CONS + tid];
                            int win = -1;
#pragma unroll
                            for (int a = 0; a < F_GW; a++)
                                if (a < na && acc[a][r] > bv) { bv = acc[a][r]; win = a; }
                            if (win >= 0) {
                                best_s[r * F_CONS + tid] = bv;
                                bestk_s[r * F_CONS + tid] = kbase + (unsigned)win;
                            }
                        }
                    } else {
                        int2 rg[F_GW];
#pragma unroll
                        for (int a = 0; a < F_GW; a++) rg[a] = rng[a];
#pragma unroll
                        for (int r = 0; r < F_R; r++) {
                            const int trel = ts + F_LANES * r;
                            float bv = best_s[r * F_CONS + tid];
                            int win = -1;
#pragma unroll
                            for (int a = 0; a < F_GW; a++) {
                                const float v = (trel >= rg[a].x && trel <= rg[a].y) ? acc[a][r] : 0.f;
                                if (a < na && v > bv) { bv = v; win = a; }
                            }
                            if (win >= 0) {
                                best_s[r * F_CONS + tid] = bv;
                                bestk_s[r * F_CONS + tid] = kbase + (unsigned)win;
                            }
                        }
                    }
                } else {
#pragma unroll
                    for (int a = 0; a < F_GW; a++) {
                        if (a < na) {
                            const size_t k = sb * F_G + j0 + half * F_GW + a;
                            const int2 rg = rng[a];
#pragma unroll
                            for (int r = 0; r < F_R; r++) {
                                const int trel = ts + F_LANES * r;
                                const float v = (trel >= rg.x && trel <= rg.y) ? acc[a][r] : 0.f;
                                if (t0 + trel < N) beam[k * (size_t)N + (size_t)(t0 + trel)] = v;
                            }
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(empty + stage);
                DIAG_ADD(d_epi, c_epi);
            }
        }
#ifdef BEAM_DIAG
        if (tid == 0) {
            atomicAdd(beam_diag + 3, (unsigned long long)d_full);
            atomicAdd(beam_diag + 4, (unsigned long long)d_epi);
            atomicAdd(beam_diag + 5, (unsigned long long)(clock64() - d_start));
        }
#endif
        if (REDUCE_MAX && any) {
#pragma unroll
            for (int r = 0; r < F_R; r++) {
                const long long t = t0 + ts + F_LANES * r;
                if (t < N) {
                    const unsigned long long key = ((unsigned long long)ordered_bits(best_s[r * F_CONS + tid]) << 32) |
                                                   (unsigned long long)(~bestk_s[r * F_CONS + tid]);
                    atomicMax(keys + t, key);
                }
            }
        }
    }
}

// Sources whose moveouts are unrelated to their neighbours' (groups of one or two): no window is shared, so
// nothing is staged -- every gather is a coalesced 128-byte read of the L2-resident row.  D_ROWS rows x D_R
// samples of loads are in flight per thread; (station, phase) rows are added in the same order as in
// beam_group_kernel, so both kernels give the same bits for the same source.
constexpr int D_THREADS = 256, D_R = 4, D_TT = D_THREADS * D_R, D_KB = 4, D_ROWS = 4;

template <bool REDUCE_MAX>
__global__ void __launch_bounds__(D_THREADS)
    beam_direct_kernel(const float *__restrict__ U, const int *__restrict__ tau, const float *__restrict__ beta,
                       const int *__restrict__ dlist, const int *__restrict__ dcount, const int *__restrict__ tlo,
                       const int *__restrict__ thi, int S, int P, long long N, long long pitch, int strict,
                       unsigned long long *__restrict__ keys, float *__restrict__ beam) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int SP = S * P;
    int *tau_s = reinterpret_cast<int *>(smem_raw);                 // D_KB x SP
    float *beta_s = reinterpret_cast<float *>(tau_s + D_KB * SP);   // D_KB x S
    int *k_s = reinterpret_cast<int *>(beta_s + D_KB * S);          // D_KB
    const int n_d = *dcount;
    const int tid = threadIdx.x;
    const long long t0 = (long long)blockIdx.x * D_TT;
    float best[D_R];
    unsigned bestk[D_R];
#pragma unroll
    for (int r = 0; r < D_R; r++) { best[r] = -INFINITY; bestk[r] = 0u; }
    bool any = false;
    for (int i0 = blockIdx.y * D_KB; i0 < n_d; i0 += gridDim.y * D_KB) {
        const int nk = (n_d - i0 < D_KB) ? (n_d - i0) : D_KB;
        __syncthreads();
        for (int e = tid; e < nk * SP; e += D_THREADS) {
            const int kk = e / SP;
            tau_s[e] = tau[(size_t)dlist[i0 + kk] * SP + (e - kk * SP)];
        }
        for (int e = tid; e < nk * S; e += D_THREADS) {
            const int kk = e / S;
            beta_s[e] = beta[(size_t)dlist[i0 + kk] * S + (e - kk * S)];
        }
        if (tid < nk) k_s[tid] = dlist[i0 + tid];
        __syncthreads();
        any = true;
        for (int kk = 0; kk < nk; kk++) {
            float acc[D_R];
#pragma unroll
            for (int r = 0; r < D_R; r++) acc[r] = 0.f;
            for (int sp0 = 0; sp0 < SP; sp0 += D_ROWS) {
                float x[D_ROWS][D_R], b[D_ROWS];
#pragma unroll
                for (int i = 0; i < D_ROWS; i++) {
                    const int sp = sp0 + i;
                    b[i] = (sp < SP) ? beta_s[kk * S + sp / P] : 0.f;
                    const long long tv = (sp < SP) ? tau_s[kk * SP + sp] : 0;
                    const float *row = U + (size_t)(sp < SP ? sp : 0) * pitch;
#pragma unroll
                    for (int r = 0; r < D_R; r++) {
                        const long long gi = t0 + tid + D_THREADS * r + tv;
                        x[i][r] = (b[i] != 0.f && gi >= 0 && gi < N) ? __ldg(row + gi) : 0.f;
                    }
                }
#pragma unroll
                for (int i = 0; i < D_ROWS; i++)
                    if (b[i] != 0.f) {
#pragma unroll
                        for (int r = 0; r < D_R; r++) acc[r] = fmaf(b[i], x[i][r], acc[r]);
                    }
            }
            const int k = k_s[kk];
            long long lo_k = 0, hi_k = N - 1;
            if (strict) { lo_k = tlo[k]; hi_k = thi[k]; }
#pragma unroll
            for (int r = 0; r < D_R; r++) {
                const long long t = t0 + tid + D_THREADS * r;
                const float v = (t >= lo_k && t <= hi_k) ? acc[r] : 0.f;
                if (REDUCE_MAX) {
                    if (v > best[r] || (v == best[r] && (unsigned)k < bestk[r])) { best[r] = v; bestk[r] = (unsigned)k; }
                } else if (t < N) {
                    beam[(size_t)k * (size_t)N + (size_t)t] = v;
                }
            }
        }
    }
    if (REDUCE_MAX && any) {
#pragma unroll
        for (int r = 0; r < D_R; r++) {
            const long long t = t0 + tid + D_THREADS * r;
            if (t < N) {
                const unsigned long long key =
                    ((unsigned long long)ordered_bits(best[r]) << 32) | (unsigned long long)(~bestk[r]);
                atomicMax(keys + t, key);
            }
        }
    }
}

static size_t beam_group_smem_bytes(size_t S, size_t P, bool reduce_max, int split) {
    return sizeof(float) * (F_NSTAGE * F_STAGE + F_TT) + 2 * (sizeof(int) * S * P * F_G + sizeof(float) * S * F_G) +
           sizeof(uint64_t) * (2 * F_NSTAGE + 2) + (reduce_max ? 2 * sizeof(float) * F_TT * (size_t)split : 0);
}

__global__ void beam_finalize_kernel(const unsigned long long *__restrict__ keys, size_t N,
                                     int source_offset, float *__restrict__ maxbeam,
                                     int *__restrict__ argmax) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N) return;
    const unsigned long long key = keys[t];
    if (key == 0ull) {  // no source at all
        maxbeam[t] = 0.f;
        argmax[t] = source_offset;
        return;
    }
    maxbeam[t] = from_ordered_bits((unsigned)(key >> 32));
    argmax[t] = (int)(~(unsigned)(key & 0xffffffffull)) + source_offset;
}

__global__ void beam_merge_kernel(float *__restrict__ ma, int *__restrict__ aa,
                                  const float *__restrict__ mb, const int *__restrict__ ab, size_t N) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N) return;
    const float va = ma[t], vb = mb[t];
    const int ka = aa[t], kb = ab[t];
    if (vb > va || (vb == va && kb < ka)) { ma[t] = vb; aa[t] = kb; }
}

static size_t beam_smem_bytes(size_t S, size_t P) {
    return sizeof(float) * B_NBUF * B_CAP + sizeof(int) * B_KB * S * P + sizeof(float) * B_KB * S +
           2 * sizeof(int) * (S * P + 1) + sizeof(int2) * B_KB * S * P + 16;
}

struct BeamWs {
    float *U;
    int *tlo, *thi;
    unsigned long long *keys;
    int *tauT;             // (n_sb, S*P, 16) transposed moveouts of the 16-source super-blocks
    float *betaT;          // (n_sb, S, 16)
    unsigned short *grp;   // (n_sb, 16) groups: start | count << 8, 0 terminates
    int *dlist, *dcount;   // sources of the groups of <= F_DIRECT_MAX sources (any order), and how many
};

static BeamWs carve(Workspace &ws, size_t K, size_t S, size_t P, size_t N) {
    BeamWs b;
    b.U = ws.take<float>(S * P * align_up(N, 4));
    b.tlo = ws.take<int>(K);
    b.thi = ws.take<int>(K);
    b.keys = ws.take<unsigned long long>(N);
    const size_t n_sb = (K + F_G - 1) / F_G;
    b.tauT = ws.take<int>(n_sb * S * P * F_G);
    b.betaT = ws.take<float>(n_sb * S * F_G);
    b.grp = ws.take<unsigned short>(n_sb * F_G);
    b.dlist = ws.take<int>(K);
    b.dcount = ws.take<int>(64);
    return b;
}

#ifndef BEAM_SPLIT_DEFAULT
#define BEAM_SPLIT_DEFAULT 2
#endif

template <bool REDUCE_MAX, int SPLIT, typename... Args>
static int launch_group_split(dim3 grid, size_t smem, cudaStream_t st, Args... args) {
    BPMF_CUDA(cudaFuncSetAttribute(beam_group_kernel<REDUCE_MAX, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem));
    beam_group_kernel<REDUCE_MAX, SPLIT><<<grid, F_THREADS, smem, st>>>(args...);
    BPMF_LAUNCH_CHECK();
    return BPMF_OK;
}

template <bool REDUCE_MAX, typename... Args>
static int launch_group(int split, dim3 grid, size_t smem, cudaStream_t st, Args... args) {
    if (split == 1) return launch_group_split<REDUCE_MAX, 1>(grid, smem, st, args...);
    return launch_group_split<REDUCE_MAX, 2>(grid, smem, st, args...);
}

static int beam_run(bool reduce_max, const float *d_features, const int *d_tau, const float *d_wp,
                    const float *d_beta, size_t K, size_t S, size_t C, size_t P, size_t N,
                    int flexible, int source_offset, void *d_ws, size_t ws_bytes, float *d_max,
                    int *d_arg, float *d_beam, cudaStream_t st) {
    BPMF_REQUIRE(S > 0 && C > 0 && P > 0 && N > 0, "beamform: empty input");
    BPMF_REQUIRE(N < (size_t)1 << 31 && K < (size_t)1 << 31, "beamform: size exceeds int32 indexing");
    Workspace ws(d_ws, ws_bytes);
    BeamWs b = carve(ws, K, S, P, N);
    BPMF_REQUIRE(ws.ok(), "beamform: workspace too small (%zu < %zu)", ws_bytes, ws.used);
    const size_t smem = beam_smem_bytes(S, P);
    const size_t pitch = align_up(N, 4);
    BPMF_REQUIRE(smem <= 200 * 1024, "beamform: too many stations x phases for one tile (%zu x %zu)", S, P);

    {
        dim3 grid((unsigned)((pitch + 1023) / 1024), (unsigned)(S * P));
        beam_prestack_kernel<<<grid, 256, 0, st>>>(d_features, d_wp, (int)C, (int)P, N, pitch, b.U);
        BPMF_LAUNCH_CHECK();
    }
    if (K > 0) {
        beam_source_bounds_kernel<<<(unsigned)((K + 127) / 128), 128, 0, st>>>(
            d_tau, d_beta, K, (int)S, (int)P, (long long)N, b.tlo, b.thi);
        BPMF_LAUNCH_CHECK();
    }
    // grouped warp-specialised kernel unless its shared-memory footprint does not fit the SM
    // (hundreds of stations) or BPMF_BEAM_KERNEL=legacy asks for the barrier-synchronised one
    int split = BEAM_SPLIT_DEFAULT;  // consumer warps per group of samples (BPMF_BEAM_SPLIT = 1 / 2 for A/B runs)
    if (const char *e = getenv("BPMF_BEAM_SPLIT")) split = atoi(e);
    split = split == 1 ? 1 : 2;
    const size_t gsmem = beam_group_smem_bytes(S, P, reduce_max, split);
    bool grouped = gsmem <= (size_t)(226 * 1024) / BEAM_MINB - 1024;  // BEAM_MINB CTAs per SM (1 KB reserved each)
    if (const char *e = getenv("BPMF_BEAM_KERNEL")) grouped = grouped && strcmp(e, "legacy") != 0;
    const size_t n_sb = (K + F_G - 1) / F_G;
    if (grouped && K > 0) {
        const int warps = 4;
        BPMF_CUDA(cudaMemsetAsync(b.dcount, 0, sizeof(int), st));
        beam_pack_kernel<<<(unsigned)((n_sb + warps - 1) / warps), warps * 32,
                           sizeof(long long) * 2 * S * P * warps, st>>>(d_tau, d_beta, K, (int)S, (int)P, b.tauT,
                                                                         b.betaT, b.grp, b.dlist, b.dcount);
        BPMF_LAUNCH_CHECK();
    }
    const size_t n_tiles = grouped ? (N + F_TT - 1) / F_TT : (N + B_TT - 1) / B_TT;
    // direct kernel: its CTAs find the number of listed sources on the device and leave when there are none
    const size_t d_tiles = (N + D_TT - 1) / D_TT;
    const size_t dsmem = sizeof(int) * D_KB * (S * P + S + 1);
    const dim3 dgrid((unsigned)d_tiles, (unsigned)std::max<size_t>(1, std::min<size_t>(64, (8 * (size_t)num_sms() + d_tiles - 1) / d_tiles)));
    if (reduce_max) {
        BPMF_CUDA(cudaMemsetAsync(b.keys, 0, sizeof(unsigned long long) * N, st));
        if (K > 0) {
            MainKernelTimer timer(st);
            // enough source chunks to fill the machine when the time axis is short
            const size_t want = (size_t)num_sms() * 4;
            size_t n_chunks = (n_tiles >= want) ? 1 : (want + n_tiles - 1) / n_tiles;
            if (grouped) {
                // CTAs run in waves of (SMs x BEAM_MINB); a short time axis gives only a few waves, so the source
                // chunking is chosen to leave the last wave as full as possible (4.56 waves -> 91 % of the machine)
                const size_t slots = (size_t)num_sms() * BEAM_MINB;
                if (n_tiles < 16 * slots) {
                    const size_t c_lo = std::max<size_t>(1, (4 * slots + n_tiles - 1) / n_tiles);
                    const size_t c_hi = std::max<size_t>(c_lo, (12 * slots + n_tiles - 1) / n_tiles);
                    double best_eff = -1.0;
                    for (size_t c = c_lo; c <= c_hi && c <= n_sb; c++) {
                        const size_t sbc = (n_sb + c - 1) / c, cc = (n_sb + sbc - 1) / sbc;
                        const size_t items = n_tiles * cc, waves = (items + slots - 1) / slots;
                        const double eff = (double)items / (double)(waves * slots);
                        if (eff > best_eff + 1e-3) { best_eff = eff; n_chunks = cc; }
                    }
                }
                size_t sb_chunk = (n_sb + n_chunks - 1) / n_chunks;
                n_chunks = (n_sb + sb_chunk - 1) / sb_chunk;
                BPMF_REQUIRE(n_chunks <= 65535, "beamform: too many source chunks");
                if (int rc = launch_group<true>(split, dim3((unsigned)n_tiles, (unsigned)n_chunks), gsmem, st, b.U, b.tauT,
                                              b.betaT, b.grp, b.tlo, b.thi, K, (int)S, (int)P, (long long)N,
                                              (long long)pitch, flexible ? 0 : 1, sb_chunk, b.keys, nullptr)) return rc;
                beam_direct_kernel<true><<<dgrid, D_THREADS, dsmem, st>>>(
                    b.U, d_tau, d_beta, b.dlist, b.dcount, b.tlo, b.thi, (int)S, (int)P, (long long)N,
                    (long long)pitch, flexible ? 0 : 1, b.keys, nullptr);
            } else {
                size_t k_chunk = align_up((K + n_chunks - 1) / n_chunks, B_KB);
                n_chunks = (K + k_chunk - 1) / k_chunk;
                BPMF_REQUIRE(n_chunks <= 65535, "beamform: too many source chunks");
                BPMF_CUDA(cudaFuncSetAttribute(beam_kernel<true>,
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                dim3 grid((unsigned)n_tiles, (unsigned)n_chunks);
                beam_kernel<true><<<grid, B_THREADS, smem, st>>>(b.U, d_tau, d_beta, b.tlo, b.thi, K,
                                                                 (int)S, (int)P, (long long)N, (long long)pitch,
                                                                 flexible ? 0 : 1, k_chunk, b.keys, nullptr);
            }
            BPMF_LAUNCH_CHECK();
        }
        beam_finalize_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(b.keys, N, source_offset,
                                                                         d_max, d_arg);
        BPMF_LAUNCH_CHECK();
    } else if (K > 0 && grouped) {
        size_t sb_chunk = 4;
        size_t n_chunks = (n_sb + sb_chunk - 1) / sb_chunk;
        if (n_chunks > 65535) {
            sb_chunk = (n_sb + 65534) / 65535;
            n_chunks = (n_sb + sb_chunk - 1) / sb_chunk;
        }
        MainKernelTimer timer(st);
        if (int rc = launch_group<false>(split, dim3((unsigned)n_tiles, (unsigned)n_chunks), gsmem, st, b.U, b.tauT, b.betaT,
                                       b.grp, b.tlo, b.thi, K, (int)S, (int)P, (long long)N, (long long)pitch,
                                       flexible ? 0 : 1, sb_chunk, nullptr, d_beam)) return rc;
        beam_direct_kernel<false><<<dgrid, D_THREADS, dsmem, st>>>(
            b.U, d_tau, d_beta, b.dlist, b.dcount, b.tlo, b.thi, (int)S, (int)P, (long long)N, (long long)pitch,
            flexible ? 0 : 1, nullptr, d_beam);
        BPMF_LAUNCH_CHECK();
    } else if (K > 0) {
        size_t k_chunk = B_KB * 8;
        size_t n_chunks = (K + k_chunk - 1) / k_chunk;
        if (n_chunks > 65535) {
            k_chunk = align_up((K + 65534) / 65535, B_KB);
            n_chunks = (K + k_chunk - 1) / k_chunk;
        }
        BPMF_CUDA(cudaFuncSetAttribute(beam_kernel<false>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        MainKernelTimer timer(st);
        dim3 grid((unsigned)n_tiles, (unsigned)n_chunks);
        beam_kernel<false><<<grid, B_THREADS, smem, st>>>(b.U, d_tau, d_beta, b.tlo, b.thi, K, (int)S,
                                                          (int)P, (long long)N, (long long)pitch, flexible ? 0 : 1,
                                                          k_chunk, nullptr, d_beam);
        BPMF_LAUNCH_CHECK();
    }
    return BPMF_OK;
}

}  // namespace bpmf

using namespace bpmf;

#ifdef BEAM_DIAG
extern "C" int bpmf_beam_diag(unsigned long long *out, int reset) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, bpmf::beam_diag, sizeof(unsigned long long) * 8);
    if (reset) {
        unsigned long long z[8] = {0};
        cudaMemcpyToSymbol(bpmf::beam_diag, z, sizeof(z));
    }
    return 0;
}
#endif

extern "C" size_t bpmf_beam_workspace_bytes(size_t K, size_t S, size_t C, size_t P, size_t N) {
    (void)C;
    Workspace ws(nullptr, 0);
    carve(ws, K, S, P, N);
    return ws.used;
}

extern "C" int bpmf_beamform_max_dev(const float *d_features, const int *d_moveouts,
                                     const float *d_weights_phases, const float *d_weights_sources,
                                     size_t K, size_t S, size_t C, size_t P, size_t N, int flexible,
                                     int source_offset, void *d_workspace, size_t workspace_bytes,
                                     float *d_maxbeam, int *d_maxbeam_sources, void *stream) {
    return beam_run(true, d_features, d_moveouts, d_weights_phases, d_weights_sources, K, S, C, P, N,
                    flexible, source_offset, d_workspace, workspace_bytes, d_maxbeam,
                    d_maxbeam_sources, nullptr, (cudaStream_t)stream);
}

extern "C" int bpmf_beamform_full_dev(const float *d_features, const int *d_moveouts,
                                      const float *d_weights_phases, const float *d_weights_sources,
                                      size_t K, size_t S, size_t C, size_t P, size_t N, int flexible,
                                      void *d_workspace, size_t workspace_bytes, float *d_beam,
                                      void *stream) {
    return beam_run(false, d_features, d_moveouts, d_weights_phases, d_weights_sources, K, S, C, P, N,
                    flexible, 0, d_workspace, workspace_bytes, nullptr, nullptr, d_beam,
                    (cudaStream_t)stream);
}

extern "C" int bpmf_beam_merge_dev(float *d_max_a, int *d_arg_a, const float *d_max_b,
                                   const int *d_arg_b, size_t n_samples, void *stream) {
    if (n_samples == 0) return BPMF_OK;
    beam_merge_kernel<<<(unsigned)((n_samples + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        d_max_a, d_arg_a, d_max_b, d_arg_b, n_samples);
    BPMF_LAUNCH_CHECK();
    return BPMF_OK;
}

// Host-pointer form (what beampower.beamform calls).  Calls whose device footprint stays under 1 GB run out of the
// calling thread's grow-only scratch and stream (HostPool: no cudaMalloc / cudaFree per call -- BPMF calls this once
// per batch of sources or per day); larger ones allocate and free their own buffers so that nothing of that size
// stays resident between calls.
static int beam_host(bool reduce_max, const float *features, const int *moveouts, const float *wp,
                     const float *beta, size_t K, size_t S, size_t C, size_t P, size_t N, int flexible,
                     int source_offset, float *maxbeam, int *arg, float *beam) {
    const size_t wsb = bpmf_beam_workspace_bytes(K, S, C, P, N);
    const size_t n_f = S * C * N, n_tau = K * S * P, n_wp = S * C * P, n_beta = K * S;
    const size_t n_out = reduce_max ? N : K * N;
    const size_t total = sizeof(float) * (n_f + n_wp + n_beta + n_out) + sizeof(int) * (n_tau + (reduce_max ? N : 0)) + wsb +
                         8 * 256;
    DevBuf f, tau, dwp, dbeta, ws, o1, o2;
    float *d_f, *d_wp, *d_beta, *d_o1;
    int *d_tau, *d_o2 = nullptr;
    void *d_ws;
    cudaStream_t st = nullptr;
    if (total <= ((size_t)1 << 30)) {
        HostPool &pool = host_pool();
        BPMF_CUDA(pool.begin(total));
        st = pool.st;
        d_f = pool.take<float>(n_f);
        d_tau = pool.take<int>(n_tau);
        d_wp = pool.take<float>(n_wp);
        d_beta = pool.take<float>(n_beta);
        d_o1 = pool.take<float>(n_out);
        if (reduce_max) d_o2 = pool.take<int>(N);
        d_ws = pool.take<char>(wsb);
    } else {
        BPMF_CUDA(f.alloc(sizeof(float) * n_f));
        BPMF_CUDA(tau.alloc(sizeof(int) * n_tau));
        BPMF_CUDA(dwp.alloc(sizeof(float) * n_wp));
        BPMF_CUDA(dbeta.alloc(sizeof(float) * n_beta));
        BPMF_CUDA(ws.alloc(wsb));
        BPMF_CUDA(o1.alloc(sizeof(float) * n_out));
        if (reduce_max) BPMF_CUDA(o2.alloc(sizeof(int) * N));
        d_f = f.as<float>();
        d_tau = tau.as<int>();
        d_wp = dwp.as<float>();
        d_beta = dbeta.as<float>();
        d_o1 = o1.as<float>();
        d_o2 = o2.as<int>();
        d_ws = ws.p;
    }
    BPMF_CUDA(cudaMemcpyAsync(d_f, features, sizeof(float) * n_f, cudaMemcpyHostToDevice, st));
    BPMF_CUDA(cudaMemcpyAsync(d_tau, moveouts, sizeof(int) * n_tau, cudaMemcpyHostToDevice, st));
    BPMF_CUDA(cudaMemcpyAsync(d_wp, wp, sizeof(float) * n_wp, cudaMemcpyHostToDevice, st));
    BPMF_CUDA(cudaMemcpyAsync(d_beta, beta, sizeof(float) * n_beta, cudaMemcpyHostToDevice, st));
    int rc;
    if (reduce_max) {
        rc = bpmf_beamform_max_dev(d_f, d_tau, d_wp, d_beta, K, S, C, P, N, flexible, source_offset, d_ws, wsb, d_o1,
                                   d_o2, st);
        if (rc != BPMF_OK) return rc;
        BPMF_CUDA(cudaMemcpyAsync(maxbeam, d_o1, sizeof(float) * N, cudaMemcpyDeviceToHost, st));
        BPMF_CUDA(cudaMemcpyAsync(arg, d_o2, sizeof(int) * N, cudaMemcpyDeviceToHost, st));
    } else {
        rc = bpmf_beamform_full_dev(d_f, d_tau, d_wp, d_beta, K, S, C, P, N, flexible, d_ws, wsb, d_o1, st);
        if (rc != BPMF_OK) return rc;
        BPMF_CUDA(cudaMemcpyAsync(beam, d_o1, sizeof(float) * K * N, cudaMemcpyDeviceToHost, st));
    }
    BPMF_CUDA(cudaStreamSynchronize(st));
    return BPMF_OK;
}

extern "C" int bpmf_beamform_max(const float *features, const int *moveouts,
                                 const float *weights_phases, const float *weights_sources, size_t K,
                                 size_t S, size_t C, size_t P, size_t N, int flexible,
                                 int source_offset, float *maxbeam, int *maxbeam_sources) {
    return beam_host(true, features, moveouts, weights_phases, weights_sources, K, S, C, P, N, flexible,
                     source_offset, maxbeam, maxbeam_sources, nullptr);
}

extern "C" int bpmf_beamform_full(const float *features, const int *moveouts,
                                  const float *weights_phases, const float *weights_sources, size_t K,
                                  size_t S, size_t C, size_t P, size_t N, int flexible, float *beam) {
    return beam_host(false, features, moveouts, weights_phases, weights_sources, K, S, C, P, N, flexible,
                     0, nullptr, nullptr, beam);
}
