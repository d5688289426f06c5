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
    extern __shared__ __align__(16) unsigned char smem_raw[];
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
};

static BeamWs carve(Workspace &ws, size_t K, size_t S, size_t P, size_t N) {
    BeamWs b;
    b.U = ws.take<float>(S * P * align_up(N, 4));
    b.tlo = ws.take<int>(K);
    b.thi = ws.take<int>(K);
    b.keys = ws.take<unsigned long long>(N);
    return b;
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
    const size_t n_tiles = (N + B_TT - 1) / B_TT;
    if (reduce_max) {
        BPMF_CUDA(cudaMemsetAsync(b.keys, 0, sizeof(unsigned long long) * N, st));
        if (K > 0) {
            MainKernelTimer timer(st);
            // enough source chunks to fill the machine when the time axis is short
            const size_t want = (size_t)num_sms() * 4;
            size_t n_chunks = (n_tiles >= want) ? 1 : (want + n_tiles - 1) / n_tiles;
            size_t k_chunk = align_up((K + n_chunks - 1) / n_chunks, B_KB);
            n_chunks = (K + k_chunk - 1) / k_chunk;
            BPMF_REQUIRE(n_chunks <= 65535, "beamform: too many source chunks");
            BPMF_CUDA(cudaFuncSetAttribute(beam_kernel<true>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            dim3 grid((unsigned)n_tiles, (unsigned)n_chunks);
            beam_kernel<true><<<grid, B_THREADS, smem, st>>>(b.U, d_tau, d_beta, b.tlo, b.thi, K,
                                                             (int)S, (int)P, (long long)N, (long long)pitch,
                                                             flexible ? 0 : 1, k_chunk, b.keys, nullptr);
            BPMF_LAUNCH_CHECK();
        }
        beam_finalize_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(b.keys, N, source_offset,
                                                                         d_max, d_arg);
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

static int beam_host(bool reduce_max, const float *features, const int *moveouts, const float *wp,
                     const float *beta, size_t K, size_t S, size_t C, size_t P, size_t N, int flexible,
                     int source_offset, float *maxbeam, int *arg, float *beam) {
    DevBuf f, tau, dwp, dbeta, ws, o1, o2;
    BPMF_CUDA(f.alloc(sizeof(float) * S * C * N));
    BPMF_CUDA(tau.alloc(sizeof(int) * K * S * P));
    BPMF_CUDA(dwp.alloc(sizeof(float) * S * C * P));
    BPMF_CUDA(dbeta.alloc(sizeof(float) * K * S));
    const size_t wsb = bpmf_beam_workspace_bytes(K, S, C, P, N);
    BPMF_CUDA(ws.alloc(wsb));
    cudaStream_t st = nullptr;
    BPMF_CUDA(cudaMemcpyAsync(f.p, features, sizeof(float) * S * C * N, cudaMemcpyHostToDevice, st));
    BPMF_CUDA(cudaMemcpyAsync(tau.p, moveouts, sizeof(int) * K * S * P, cudaMemcpyHostToDevice, st));
    BPMF_CUDA(cudaMemcpyAsync(dwp.p, wp, sizeof(float) * S * C * P, cudaMemcpyHostToDevice, st));
    BPMF_CUDA(cudaMemcpyAsync(dbeta.p, beta, sizeof(float) * K * S, cudaMemcpyHostToDevice, st));
    int rc;
    if (reduce_max) {
        BPMF_CUDA(o1.alloc(sizeof(float) * N));
        BPMF_CUDA(o2.alloc(sizeof(int) * N));
        rc = bpmf_beamform_max_dev(f.as<float>(), tau.as<int>(), dwp.as<float>(), dbeta.as<float>(), K,
                                   S, C, P, N, flexible, source_offset, ws.p, wsb, o1.as<float>(),
                                   o2.as<int>(), st);
        if (rc != BPMF_OK) return rc;
        BPMF_CUDA(cudaMemcpyAsync(maxbeam, o1.p, sizeof(float) * N, cudaMemcpyDeviceToHost, st));
        BPMF_CUDA(cudaMemcpyAsync(arg, o2.p, sizeof(int) * N, cudaMemcpyDeviceToHost, st));
    } else {
        BPMF_CUDA(o1.alloc(sizeof(float) * K * N));
        rc = bpmf_beamform_full_dev(f.as<float>(), tau.as<int>(), dwp.as<float>(), dbeta.as<float>(), K,
                                    S, C, P, N, flexible, ws.p, wsb, o1.as<float>(), st);
        if (rc != BPMF_OK) return rc;
        BPMF_CUDA(cudaMemcpyAsync(beam, o1.p, sizeof(float) * K * N, cudaMemcpyDeviceToHost, st));
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
