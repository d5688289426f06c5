// mf_tc.cu -- matched filter on the 5th-generation tensor cores (tcgen05 / TMEM), sm_100a only.
//
// Replaces the native kernel behind fast_matched_filter.matched_filter for BPMF's default call
// (BPMF/similarity_search.py:526-533: step 1, "short" normalisation, network sum); SURVEY.md A.1.
//
// Formulation (per channel ch): the CC numerator is a Toeplitz product.  It is cast as a GEMM whose
// B operand is the RAW trace read through an overlapping shared-memory descriptor:
//
//     C[(t,r), jj] = sum_k' A[(t,r), k'] * d[8*jj + k']
//
//   * fp16 operands (per-channel / per-template power-of-two pre-scaling, exact), fp32 accumulate.
//   * B "row" jj starts 8 elements = 16 bytes after row jj-1, which is exactly the row pitch of a
//     no-swizzle K-major core matrix: the UMMA descriptor (LBO = 16 B, SBO = 128 B) walks a Hankel
//     view of the trace without any im2col copy.  One 5.7 KB slice of trace feeds a 128 x 256 tile.
//   * the 8 sub-row alignments ("phases") of a template are 8 ROWS of A: row (t, r) holds template
//     t delayed by e = r + (moveout mod 8) samples, so that row produces outputs i = r (mod 8) and
//     the 8 lanes x 1 column of a template form ONE aligned 32-byte sector of cc_sum.  The
//     whole-column part of the moveout, q = floor(moveout / 8), is an output address shift.
//   * A (16 templates x 8 phases x Kc, up to 204 KB) stays RESIDENT in shared memory while the CTA
//     sweeps a segment of 32 tiles (65 536 samples); B slices stream through a 2-stage bulk-copy
//     ring.  Templates longer than one resident image are cut into K chunks (numerators add).
//   * BPMF_MF_TC_FP16X3: both operands are split hi + lo (two fp16 = 22 mantissa bits) and three
//     MMAs (hi*hi + hi*lo + lo*hi) accumulate into the same TMEM tile: fp32-grade numerators.
//   * epilogue (8 warps): tcgen05.ld -> * rsqrt(window energy)[column] (32-column slices staged in
//     shared memory by a dedicated bulk-copy warp) * weight/|T|[row] -> red.global.add.f32 into cc_sum.  Work is ordered segment -> template tile -> channel so the
//     48 channel-CTAs that add into the same 16 rows x 131 072 samples do so while that region
//     (8 MB) is L2 resident.
//
// Rooflines: tensor pipe (2*Kp flop per output, Kp = L+7 rounded to 16) and L2 reduction
// throughput (one fp32 red per (template, channel, output); ~850 Gred/s measured, tools/red_bench.cu).
#include <cuda_fp16.h>
#include <stdlib.h>

#include "async.cuh"
#include "common.cuh"

namespace bpmf {

constexpr int TC_ROWS = 128;        // A rows per tile = 16 templates x 8 phases
constexpr int TC_TPL = 16;          // templates per tile
constexpr int TC_NCOL = 256;        // GEMM columns per tile (= 2048 samples)
constexpr int TC_KC_MAX = 816;      // resident A, single pass: 128 x 816 fp16 = 208 896 B
constexpr int TC_KC_MAX_SPLIT = 384; // resident A, hi + lo images: 128 x 384 x 2 x 2 = 196 608 B
constexpr int TC_SEG_TILES = 32;    // tiles swept per resident A image (65 536 samples)
#ifndef TC_EPI_WARPS
#define TC_EPI_WARPS 16             // 8 or 16: with 16, two warps share each (lane quarter, tile half) and take alternate chunks
#endif
constexpr int TC_EPI_WARP0 = 2;
constexpr int TC_THREADS = 32 * (TC_EPI_WARP0 + TC_EPI_WARPS + 1);  // warp 0 A/B producer, warp 1 MMA, epilogue warps, rsqrt producer
constexpr int TC_EPI_SUBS = TC_EPI_WARPS / 8;
constexpr int TC_NB = 2;            // trace-slice (B operand) stages
constexpr int TC_NI = 5;            // rsqrt(energy) slots per tile half (32 columns each)
constexpr int TC_INV_SLOT = 264;    // floats per slot: 8 * 32 columns + 8 (row delays reach 14)
constexpr int TC_TMEM_COLS = 512;
constexpr int TC_FRONT = 8;         // zero samples in front of every staged trace: column jj starts at
                                    // sample 8*(jj-1), so rows delayed by e >= 8 still reach samples 0..7

constexpr size_t TC_SMEM_LIMIT = 232448;  // 227 KB opt-in shared memory per CTA

struct TcParams {
    const __half *dh;        // (SC, dh_pitch) scaled fp16 traces (hi part), zero padded
    const __half *dh_lo;     // residual x*s - hi (split mode only)
    size_t dh_pitch;
    const float *colinv;     // (SC, ci_pitch) rsqrt(window energy) per window start, 0 where undefined
    size_t ci_pitch;
    const uint8_t *aimg;     // (SC, MT, n_kc) A images in UMMA core-matrix order, a_bytes each
                             // (split mode: hi image followed by lo image)
    const int *row_q;        // (SC, MT, 128) whole-column shift floor(moveout / 8), same for a template's 8 rows
    const int *row_e;        // (SC, MT, 128) delay folded into the A row: phase + moveout mod 8, in [0, 15)
    const float *row_scale;  // (SC, MT, 128) weight / (sA * sB * sqrt(sum T^2)), 0 = unused row
    const float *row_ithr;   // (SC, MT, 128) guard: contribute only where colinv < ithr
    const int *row_t;        // (SC, MT, 128) template of the row, -1 = empty slot (tiles are formed PER CHANNEL from
                             // the templates that weight it, so zero-weight (template, channel) pairs cost nothing)
    const int *item_active;  // (SC, MT)
    const int *lo, *hi;      // per template valid output range
    float *cc;
    size_t cc_pitch;
    int T, SC, MT, n_kc, Kc, split, a_bytes, a_img_bytes, b_slice_bytes, b_stage_bytes;
    int n_tiles, seg_tiles, n_seg;
    int debug;  // BPMF_TC_DEBUG bit 0: skip the reductions (perf triage only)
};

// A row (= TMEM lane) <-> (template of the tile, phase).
//   EPI 0: row = 8 * template + phase.  The epilogue reads one row per thread (tcgen05.ld.32x32b) and
//          transposes 4 x 4 blocks with warp shuffles to form 16-byte reductions.
//   EPI 1: within a quarter of 32 lanes, lane = 16*h + 8*v + 4*a + p holds template 4*quarter + 2*h + a,
//          phase 2*p + v: lanes l and l + 8 of a 16-lane half are the two phases of a pair.  tcgen05.ld.16x128b
//          hands thread t rows t/4 and t/4 + 8 of the half -- two CONSECUTIVE outputs of one template -- at
//          columns (t & 3) + 4k: no shuffle, one 8-byte red.v2 per column, and the 16 threads of a template
//          cover 4 columns x 8 phases = one 128-byte run per warp instruction.
__host__ __device__ __forceinline__ void tc_row_decode(int row, int epi, int &tpl, int &phase) {
    if (epi == 0) {
        tpl = row >> 3;
        phase = row & 7;
    } else {
        const int lam = row & 31;
        tpl = 4 * (row >> 5) + 2 * (lam >> 4) + ((lam >> 2) & 1);
        phase = 2 * (lam & 3) + ((lam >> 3) & 1);
    }
}

struct TcItem {
    int kc, ch, mt, seg;
};
__device__ __forceinline__ TcItem tc_decode(const TcParams &p, int it) {
    TcItem r;
    r.kc = it % p.n_kc;
    it /= p.n_kc;
    r.ch = it % p.SC;
    it /= p.SC;
    r.mt = it % p.MT;
    r.seg = it / p.MT;
    return r;
}

// ---------------------------------------------------------------- PTX wrappers (mbarrier / bulk copy: async.cuh)
__device__ __forceinline__ void red_add_f32(float *addr, float v, uint64_t policy) {
    asm volatile("red.global.add.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(addr), "f"(v), "l"(policy) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], fp16 inputs, fp32 accumulate, issued by ONE thread
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                           uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// shared-memory matrix descriptor, no swizzle, K-major: core matrix = 8 rows x 16 bytes (rows 16 B
// apart); LBO = byte step between the two 8-element K chunks of one MMA, SBO = byte step between
// consecutive 8-row groups.
__device__ __forceinline__ uint64_t tc_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version of sm_100
    return d;
}
// instruction descriptor: D = f32, A = B = f16, both K-major, M = 128, N = 256
__device__ __forceinline__ uint32_t tc_idesc(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; i++) v[i] = __uint_as_float(r[i]);
}

// 4x4 transpose between two lane bits (masks HI, LO) and the index within a group of four registers,
// by two butterfly exchanges: afterwards register x of the lane whose two bits read y holds what
// register y of the lane whose two bits read x held.
template <int HI, int LO>
__device__ __forceinline__ void tc_transpose4(float &v0, float &v1, float &v2, float &v3, bool bhi, bool blo) {
    {   // exchange across lane bit HI <-> register bit 1: pairs (0,2) and (1,3)
        float s0 = bhi ? v0 : v2, s1 = bhi ? v1 : v3;
        s0 = __shfl_xor_sync(0xffffffffu, s0, HI);
        s1 = __shfl_xor_sync(0xffffffffu, s1, HI);
        if (bhi) { v0 = s0; v1 = s1; } else { v2 = s0; v3 = s1; }
    }
    {   // exchange across lane bit LO <-> register bit 0: pairs (0,1) and (2,3)
        float s0 = blo ? v0 : v1, s1 = blo ? v2 : v3;
        s0 = __shfl_xor_sync(0xffffffffu, s0, LO);
        s1 = __shfl_xor_sync(0xffffffffu, s1, LO);
        if (blo) { v0 = s0; v2 = s1; } else { v1 = s0; v3 = s1; }
    }
}

__device__ __forceinline__ void red_add_v4(float *addr, float a, float b, float c, float d, uint64_t policy) {
    asm volatile("red.global.add.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(addr), "f"(a), "f"(b),
                 "f"(c), "f"(d), "l"(policy)
                 : "memory");
}

__device__ __forceinline__ void red_add_v2(float *addr, float a, float b, uint64_t policy) {
    asm volatile("red.global.add.L2::cache_hint.v2.f32 [%0], {%1, %2}, %3;" ::"l"(addr), "f"(a), "f"(b), "l"(policy)
                 : "memory");
}
// 16 lanes x 32 columns: thread t gets rows t/4 (even registers) and t/4 + 8 (odd registers) of the 16-lane
// group at columns (t & 3) + 4k, k = register / 2.  No wait: the caller issues tcgen05.wait::ld.
__device__ __forceinline__ void tc_ld16x128_x8(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.16x128b.x8.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}

// per-thread constants of one template-phase pair of the shuffle-free epilogue (EPI 1)
struct TcPair {
    float scale, ithr;
    float *obase;   // cc row of the template + ioff_out
    int q, ioff;    // whole-column shift; 8 * (lane & 3) + delay of phase 2p (index into the rsqrt slice)
    int ioff_out;   // 8 * (lane & 3) + 2p: the thread's first output relative to (column 0 of the chunk, phase 0)
    int lo, hi;     // valid output range of the template
    int jj_full_lo, jj_full_hi;  // columns where BOTH phases are valid
    int jj_any_lo, jj_any_hi;    // columns where at least one is
};

__device__ __forceinline__ int floor_div8(int a) { return a >> 3; }          // arithmetic shift = floor
__device__ __forceinline__ int ceil_div8(int a) { return (a + 7) >> 3; }

// one 16-lane half of a 32-column chunk: 8 columns x 2 phases per thread
__device__ __forceinline__ void tc_epi_half2(const uint32_t (&r)[16], const float *inv, const TcPair &pr, long long o,
                                             int mode, uint64_t policy) {
    // o = output index of (column 0 of the chunk, phase 0) for this template
    const float *iv = inv + pr.ioff;
    float *base = pr.obase + o;
    if (mode == 2) {
        if (pr.hi < pr.lo) return;  // unweighted template of a tile whose other templates are fully valid
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const float w0 = iv[32 * k], w1 = iv[32 * k + 1];
            const float x0 = __uint_as_float(r[2 * k]) * w0, x1 = __uint_as_float(r[2 * k + 1]) * w1;
            red_add_v2(base + 32 * k, (w0 < pr.ithr) ? x0 * pr.scale : 0.f, (w1 < pr.ithr) ? x1 * pr.scale : 0.f,
                       policy);
        }
    } else {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const float w0 = iv[32 * k], w1 = iv[32 * k + 1];
            const float x0 = __uint_as_float(r[2 * k]) * w0, x1 = __uint_as_float(r[2 * k + 1]) * w1;
            const float y0 = (w0 < pr.ithr) ? x0 * pr.scale : 0.f, y1 = (w1 < pr.ithr) ? x1 * pr.scale : 0.f;
            const long long i = o + pr.ioff_out + 32 * k;
            if (i >= pr.lo && i <= pr.hi) red_add_f32(base + 32 * k, y0, policy);
            if (i + 1 >= pr.lo && i + 1 <= pr.hi) red_add_f32(base + 32 * k + 1, y1, policy);
        }
    }
}

// one 32-column chunk of the epilogue: TMEM -> registers -> normalise -> transpose -> red.global.add.
// inv points at this row's first rsqrt(energy) value of the chunk (column c lives 8*c floats on).
// Lane (a, r) [a = lane >> 3: template of the warp, r = lane & 7: phase] holds columns c = 0..31 of its
// row; output index of (a, c, r) = 8*(column - q_a) + r, so a template's 32 columns x 8 phases are 256
// consecutive floats of its cc_sum row.  The transpose swaps the two low phase bits of the lane with the
// two low column bits of the register index: afterwards registers 4g..4g+3 of a lane are FOUR CONSECUTIVE
// outputs (one 16-byte red.v4), the 8 lanes of a template cover one 128-byte run per g, and the lane
// keeps its own template -- 8 vector reductions per chunk instead of 32 scalar ones.
// base points at output 8*(first column of the chunk - q) + 8*(lane & 3) + 4*((lane >> 2) & 1) of the lane's
// template, i0 is that output's index (valid iff lo <= i <= hi).
// mode 2: everything valid; 1: per-element predicate; 0: nothing to write.
__device__ __forceinline__ void tc_epi_chunk(uint32_t taddr, const float *inv, float *base, int i0, int lo, int hi,
                                             int mode, float ithr, float scale, bool b2, bool b1, uint64_t policy,
                                             int debug) {
    float v[32];
    tc_ld32(taddr, v);
    if (debug & 1) mode = (v[0] == 12345.678f) ? mode : 0;  // triage: keep everything but the epilogue math + reductions
    if (mode == 0) return;
#pragma unroll
    for (int c = 0; c < 32; c++) {
        const float w = inv[8 * c];
        const float x = v[c] * w;
        v[c] = (w < ithr) ? x * scale : 0.f;
    }
#pragma unroll
    for (int g = 0; g < 8; g++) tc_transpose4<2, 1>(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3], b2, b1);
    if (debug & 8) {  // triage: epilogue math without the global writes
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < 32; c++) s += v[c];
        if (s != 12345.678f) return;
    }
    if (mode == 2 && (debug & 4)) {  // triage: plain stores instead of reductions (wrong results)
#pragma unroll
        for (int g = 0; g < 8; g++)
            *reinterpret_cast<float4 *>(base + 32 * g) = make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
    } else if (mode == 2) {
#pragma unroll
        for (int g = 0; g < 8; g++) red_add_v4(base + 32 * g, v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3], policy);
    } else {
#pragma unroll
        for (int g = 0; g < 8; g++) {
            const int i = i0 + 32 * g;
            if (i >= lo && i + 3 <= hi) {
                red_add_v4(base + 32 * g, v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3], policy);
            } else {
#pragma unroll
                for (int x = 0; x < 4; x++)
                    if (i + x >= lo && i + x <= hi) red_add_f32(base + 32 * g + x, v[4 * g + x], policy);
            }
        }
    }
}

// ---------------------------------------------------------------- the kernel
template <int EPI>
__global__ void __launch_bounds__(TC_THREADS, 1) mf_tc_kernel(const TcParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *a_s = smem;
    uint8_t *b_s = smem + p.a_bytes;
    float *inv_s = reinterpret_cast<float *>(b_s + TC_NB * p.b_stage_bytes);  // [2 halves][TC_NI][TC_INV_SLOT]
    uint64_t *bars = reinterpret_cast<uint64_t *>(inv_s + 2 * TC_NI * TC_INV_SLOT);
    uint64_t *a_full = bars + 0, *a_empty = bars + 1;
    uint64_t *acc_full = bars + 2, *acc_empty = bars + 4;               // [2]
    uint64_t *b_full = bars + 6, *b_empty = bars + 6 + TC_NB;           // [TC_NB]
    uint64_t *inv_full = bars + 6 + 2 * TC_NB;                          // [2][TC_NI]
    uint64_t *inv_empty = inv_full + 2 * TC_NI;                         // [2][TC_NI]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(inv_empty + 2 * TC_NI);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        mbar_init(a_full, 1);
        mbar_init(a_empty, 1);
        for (int i = 0; i < 2; i++) {
            mbar_init(acc_full + i, 1);
            mbar_init(acc_empty + i, TC_EPI_WARPS);  // one arrive per epilogue warp
        }
        for (int i = 0; i < TC_NB; i++) {
            mbar_init(b_full + i, 1);
            mbar_init(b_empty + i, 1);
        }
        for (int i = 0; i < 2 * TC_NI; i++) {
            mbar_init(inv_full + i, 1);
            mbar_init(inv_empty + i, 4);  // the four warps that read this half
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(TC_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int n_items = p.n_seg * p.MT * p.SC * p.n_kc;
    const uint32_t b_bytes = (uint32_t)p.b_slice_bytes;  // (8*255 + Kc) * 2, a multiple of 16

    if (warp == 0) {
        // ===================== producer: bulk copies of A images and trace slices ==================
        if (lane == 0) {
            // The A images are read once per item (evict first); the trace slices and rsqrt(energy) slices of a
            // segment are read again by the next template tiles a fraction of a millisecond later: default policy
            // (measured: 168 GB of DRAM traffic per configs[1] pass against 263 GB with evict_first, same run time).
            // BPMF_TC_DEBUG bits 6/7 (tuning): evict_first / evict_last instead.
            const uint64_t pol_a = l2_policy_evict_first();
            const uint64_t pol = (p.debug & 64) ? l2_policy_evict_first()
                                                : ((p.debug & 128) ? l2_policy_evict_last() : l2_policy_evict_normal());
            uint32_t n_item = 0, n_tile = 0;
            for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
                const TcItem im = tc_decode(p, it);
                const int ch = im.ch, mt = im.mt, seg = im.seg;
                if (!p.item_active[ch * p.MT + mt]) continue;
                mbar_wait(a_empty, (n_item & 1) ^ 1);
                mbar_expect_tx(a_full, (uint32_t)p.a_bytes);
                const uint8_t *src = p.aimg + ((size_t)(ch * p.MT + mt) * p.n_kc + im.kc) * p.a_bytes;
                for (int off = 0; off < p.a_bytes; off += 16384) {
                    const int n = (p.a_bytes - off < 16384) ? (p.a_bytes - off) : 16384;
                    bulk_g2s(a_s + off, src + off, (uint32_t)n, a_full, pol_a);
                }
                n_item++;
                const int t_begin = seg * p.seg_tiles;
                const int t_end = (t_begin + p.seg_tiles < p.n_tiles) ? t_begin + p.seg_tiles : p.n_tiles;
                for (int tile = t_begin; tile < t_end; tile++, n_tile++) {
                    const int s = n_tile % TC_NB;
                    mbar_wait(b_empty + s, ((n_tile / TC_NB) & 1) ^ 1);
                    const size_t d_off = (size_t)ch * p.dh_pitch + (size_t)tile * (8 * TC_NCOL) + (size_t)im.kc * p.Kc;
                    mbar_expect_tx(b_full + s, p.split ? 2 * b_bytes : b_bytes);
                    bulk_g2s(b_s + s * p.b_stage_bytes, p.dh + d_off, b_bytes, b_full + s, pol);
                    if (p.split) bulk_g2s(b_s + s * p.b_stage_bytes + b_bytes, p.dh_lo + d_off, b_bytes, b_full + s, pol);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (one thread) ============================================
        if (lane == 0) {
            const uint32_t idesc = tc_idesc(TC_ROWS, TC_NCOL);
            const uint32_t a_addr = smem_u32(a_s), b_addr = smem_u32(b_s);
            const int n_k = p.Kc / 16;
            uint32_t n_item = 0, n_tile = 0;
            for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
                const TcItem im = tc_decode(p, it);
                const int ch = im.ch, mt = im.mt, seg = im.seg;
                if (!p.item_active[ch * p.MT + mt]) continue;
                mbar_wait(a_full, n_item & 1);
                n_item++;
                const int t_begin = seg * p.seg_tiles;
                const int t_end = (t_begin + p.seg_tiles < p.n_tiles) ? t_begin + p.seg_tiles : p.n_tiles;
                for (int k = t_begin; k < t_end; k++, n_tile++) {
                    const int s = n_tile % TC_NB;      // trace-slice stage
                    const int sa = n_tile & 1;         // accumulator stage
                    mbar_wait(b_full + s, (n_tile / TC_NB) & 1);
                    mbar_wait(acc_empty + sa, ((n_tile >> 1) & 1) ^ 1);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)(sa * TC_NCOL);
                    for (int kk = 0; kk < n_k; kk++) {
                        // A: K chunk c of 8 elements lives at c * (128 rows * 16 B); rows 16 B apart
                        const uint64_t da = tc_desc(a_addr + (uint32_t)kk * 2 * (TC_ROWS * 16), TC_ROWS * 16, 128);
                        // B: Hankel view of the raw trace: row jj at +16 B, K chunk at +16 B
                        const uint32_t b_at = b_addr + (uint32_t)(s * p.b_stage_bytes) + (uint32_t)kk * 32;
                        const uint64_t db = tc_desc(b_at, 16, 128);
                        tc_mma_f16(d_tmem, da, db, idesc, kk > 0 ? 1u : 0u);
                        if (p.split) {  // + hi*lo + lo*hi into the same accumulator
                            const uint64_t db_lo = tc_desc(b_at + b_bytes, 16, 128);
                            const uint64_t da_lo = tc_desc(a_addr + (uint32_t)p.a_img_bytes + (uint32_t)kk * 2 * (TC_ROWS * 16),
                                                           TC_ROWS * 16, 128);
                            tc_mma_f16(d_tmem, da, db_lo, idesc, 1u);
                            tc_mma_f16(d_tmem, da_lo, db, idesc, 1u);
                        }
                    }
                    tc_commit(b_empty + s);    // trace slice may be overwritten once these MMAs retire
                    tc_commit(acc_full + sa);  // accumulator ready for the epilogue
                }
                tc_commit(a_empty);  // A image may be replaced once every MMA of this item retired
            }
        }
    } else if (warp == TC_EPI_WARP0 + TC_EPI_WARPS) {
        // ===================== rsqrt(energy) producer: 32-column slices for the epilogue ===========
        if (lane == 0) {
            const uint64_t pol = (p.debug & 64) ? l2_policy_evict_first()
                                                : ((p.debug & 128) ? l2_policy_evict_last() : l2_policy_evict_normal());
            uint32_t n_chunk = 0;  // per half: chunks issued so far
            for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
                const TcItem im = tc_decode(p, it);
                const int ch = im.ch, mt = im.mt, seg = im.seg;
                if (!p.item_active[ch * p.MT + mt]) continue;
                const int t_begin = seg * p.seg_tiles;
                const int t_end = (t_begin + p.seg_tiles < p.n_tiles) ? t_begin + p.seg_tiles : p.n_tiles;
                const float *ci = p.colinv + (size_t)ch * p.ci_pitch;
                for (int tile = t_begin; tile < t_end; tile++) {
                    for (int chunk = 0; chunk < 4; chunk++, n_chunk++) {
                        const int slot = n_chunk % TC_NI;
                        const uint32_t par = ((n_chunk / TC_NI) & 1) ^ 1;
#pragma unroll
                        for (int h = 0; h < 2; h++) {
                            uint64_t *full = inv_full + h * TC_NI + slot;
                            mbar_wait(inv_empty + h * TC_NI + slot, par);
                            mbar_expect_tx(full, TC_INV_SLOT * 4);
                            bulk_g2s(inv_s + (h * TC_NI + slot) * TC_INV_SLOT,
                                     ci + 8 * ((size_t)tile * TC_NCOL + h * 128 + chunk * 32), TC_INV_SLOT * 4, full, pol);
                        }
                    }
                }
            }
        }
    } else if (EPI == 1) {
        // ===================== epilogue, shuffle-free: 16x128b TMEM loads -> normalise -> red.v2 ==========
        const int quarter = warp & 3;                      // TMEM lane quarter this warp may read
        const int half = ((warp - TC_EPI_WARP0) >> 2) & 1; // which 128 columns of the tile
        const int sub = (warp - TC_EPI_WARP0) >> 3;        // which chunks of that half: sub, sub + TC_EPI_SUBS, ...
        const uint64_t red_policy = (p.debug & 256) ? l2_policy_evict_normal() : l2_policy_evict_last();
        const int pp = (lane >> 2) & 3;                    // phase pair of this thread: phases 2pp, 2pp + 1
        uint32_t n_tile = 0;
        for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
            const TcItem im = tc_decode(p, it);
            const int ch = im.ch, mt = im.mt, seg = im.seg;
            if (!p.item_active[ch * p.MT + mt]) continue;
            TcPair pr[2];
#pragma unroll
            for (int hh = 0; hh < 2; hh++) {
                // row of phase 2pp of this thread's template in 16-lane half hh (lam = 16 hh + 4 a + pp)
                const int row0 = quarter * 32 + 16 * hh + 4 * (lane >> 4) + pp;
                const size_t mrow = (size_t)(ch * p.MT + mt) * TC_ROWS + row0;
                const int t = p.row_t[mrow];
                TcPair &r = pr[hh];
                r.scale = p.row_scale[mrow];
                r.ithr = p.row_ithr[mrow];
                r.q = p.row_q[mrow];
                r.ioff = 8 * (lane & 3) + p.row_e[mrow];
                r.ioff_out = 8 * (lane & 3) + 2 * pp;
                r.lo = 1;
                r.hi = 0;
                if (t >= 0 && r.scale != 0.f) { r.lo = p.lo[t]; r.hi = p.hi[t]; }
                r.obase = p.cc + (size_t)(t >= 0 ? t : 0) * p.cc_pitch + r.ioff_out;
                r.jj_full_lo = r.jj_any_lo = 1;
                r.jj_full_hi = r.jj_any_hi = 0;
                if (r.hi >= r.lo) {  // output of (column jj, phase ph) = 8 * (jj - q) + ph
                    r.jj_full_lo = ceil_div8(r.lo - 2 * pp) + r.q;
                    r.jj_full_hi = floor_div8(r.hi - 2 * pp - 1) + r.q;
                    r.jj_any_lo = ceil_div8(r.lo - 2 * pp - 1) + r.q;
                    r.jj_any_hi = floor_div8(r.hi - 2 * pp) + r.q;
                }
            }
            const int t_begin = seg * p.seg_tiles;
            const int t_end = (t_begin + p.seg_tiles < p.n_tiles) ? t_begin + p.seg_tiles : p.n_tiles;
            for (int tile = t_begin; tile < t_end; tile++, n_tile++) {
                const int s = n_tile & 1;
                mbar_wait(acc_full + s, (n_tile >> 1) & 1);
                tc_fence_after();
                const int jj0 = tile * TC_NCOL + half * 128;  // first column of this half
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(s * TC_NCOL + half * 128);
                int mode[2];
#pragma unroll
                for (int hh = 0; hh < 2; hh++) {
                    const bool full = (pr[hh].hi < pr[hh].lo) || (pr[hh].jj_full_lo <= jj0 && pr[hh].jj_full_hi >= jj0 + 127);
                    const bool any = pr[hh].jj_any_lo <= jj0 + 127 && pr[hh].jj_any_hi >= jj0;
                    mode[hh] = !__any_sync(0xffffffffu, any) ? 0 : (__all_sync(0xffffffffu, full) ? 2 : 1);
                }
                if (p.debug & 1) mode[0] = mode[1] = 0;  // triage: MMA + producers only
#pragma unroll 1
                for (int chunk = sub; chunk < 4; chunk += TC_EPI_SUBS) {
                    const uint32_t n_chunk = 4 * n_tile + (uint32_t)chunk;  // chunks of this half issued so far
                    const int slot = n_chunk % TC_NI;
                    uint64_t *full = inv_full + half * TC_NI + slot;
                    uint32_t r0[16], r1[16];
                    tc_ld16x128_x8(taddr + (uint32_t)(chunk * 32), r0);
                    tc_ld16x128_x8(taddr + (16u << 16) + (uint32_t)(chunk * 32), r1);
                    mbar_wait(full, (n_chunk / TC_NI) & 1);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    const float *inv = inv_s + (half * TC_NI + slot) * TC_INV_SLOT;
                    if (mode[0]) tc_epi_half2(r0, inv, pr[0], 8ll * (jj0 + chunk * 32 - pr[0].q), mode[0], red_policy);
                    if (mode[1]) tc_epi_half2(r1, inv, pr[1], 8ll * (jj0 + chunk * 32 - pr[1].q), mode[1], red_policy);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(inv_empty + half * TC_NI + slot);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(acc_empty + s);
            }
        }
    } else {
        // ===================== epilogue: TMEM -> normalise -> red.add into cc_sum ==================
        const int quarter = warp & 3;                      // TMEM lane quarter this warp may read
        const int half = ((warp - TC_EPI_WARP0) >> 2) & 1; // which 128 columns of the tile
        const int sub = (warp - TC_EPI_WARP0) >> 3;        // which chunks of that half: sub, sub + TC_EPI_SUBS, ...
        const int row = quarter * 32 + lane;               // A row == TMEM lane
        const uint64_t red_policy = (p.debug & 32) ? l2_policy_evict_first() : l2_policy_evict_last();
        uint32_t n_tile = 0;
        for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
            const TcItem im = tc_decode(p, it);
            const int ch = im.ch, mt = im.mt, seg = im.seg;
            if (!p.item_active[ch * p.MT + mt]) continue;
            const size_t mrow = (size_t)(ch * p.MT + mt) * TC_ROWS + row;
            const float scale = p.row_scale[mrow];
            const float ithr = p.row_ithr[mrow];
            const int q = p.row_q[mrow];
            const int e = p.row_e[mrow];  // in [0, 15)
            const int t = p.row_t[mrow];
            const int ir = row & 7;
            int lo = 1, hi = 0;
            if (t >= 0 && scale != 0.f) { lo = p.lo[t]; hi = p.hi[t]; }
            // output index of column jj: i = 8 * (jj - q) + ir, valid iff lo <= i <= hi
            // <=> jj in [jj_lo, jj_hi]
            long long jj_lo = 1, jj_hi = 0;
            if (hi >= lo) {
                const int a = lo - ir, b = hi - ir;
                jj_lo = (long long)((a >= 0) ? (a + 7) / 8 : -((-a) / 8)) + q;  // ceil(a / 8) + q
                jj_hi = (long long)((b >= 0) ? b / 8 : -((-b + 7) / 8)) + q;    // floor(b / 8) + q
            }
            // after the transpose the lane still writes its own template: 16-byte pieces at
            // 8 * (lane & 3) + 4 * ((lane >> 2) & 1) floats into each 128-byte run
            float *obase = p.cc + (size_t)(t >= 0 ? t : 0) * p.cc_pitch + 8 * (lane & 3) + 4 * ((lane >> 2) & 1);
            const bool b2 = (lane & 2) != 0, b1 = (lane & 1) != 0;
            const int t_begin = seg * p.seg_tiles;
            const int t_end = (t_begin + p.seg_tiles < p.n_tiles) ? t_begin + p.seg_tiles : p.n_tiles;
            for (int tile = t_begin; tile < t_end; tile++, n_tile++) {
                const int s = n_tile & 1;
                mbar_wait(acc_full + s, (n_tile >> 1) & 1);
                tc_fence_after();
                const long long jj0 = (long long)tile * TC_NCOL + half * 128;  // first column of this half
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(s * TC_NCOL + half * 128);
                // columns of this half that are valid for this row: [c_lo, c_hi)
                long long c_lo64 = jj_lo - jj0, c_hi64 = jj_hi - jj0 + 1;
                const int c_lo = (int)(c_lo64 < 0 ? 0 : (c_lo64 > 128 ? 128 : c_lo64));
                const int c_hi = (int)(c_hi64 > 128 ? 128 : (c_hi64 < 0 ? 0 : c_hi64));
                const int mode = __all_sync(0xffffffffu, c_lo == 0 && c_hi == 128)
                                     ? 2 : (__any_sync(0xffffffffu, c_hi > c_lo) ? 1 : 0);
#pragma unroll 1
                for (int chunk = sub; chunk < 4; chunk += TC_EPI_SUBS) {
                    const uint32_t n_chunk = 4 * n_tile + (uint32_t)chunk;  // chunks of this half issued so far
                    const int slot = n_chunk % TC_NI;
                    uint64_t *full = inv_full + half * TC_NI + slot;
                    mbar_wait(full, (n_chunk / TC_NI) & 1);
                    const float *inv = inv_s + (half * TC_NI + slot) * TC_INV_SLOT + e;
                    const long long o = 8 * (jj0 + chunk * 32 - q);  // output index of column 0, phase 0
                    tc_epi_chunk(taddr + (uint32_t)(chunk * 32), inv, obase + o,
                                 (int)(o + 8 * (lane & 3) + 4 * ((lane >> 2) & 1)), lo, hi, mode, ithr, scale, b2, b1,
                                 red_policy, p.debug);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(inv_empty + half * TC_NI + slot);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(acc_empty + s);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS)
                     : "memory");
    }
}

// ---------------------------------------------------------------- the kernel, CTA pairs (cta_group::2)
// Two CTAs of a cluster (the two SMs of a TPC) work on the SAME channel and the SAME trace tiles with two different
// template tiles: one tcgen05.mma.cta_group::2 of M = 256 covers both, each CTA keeping its own resident A image and
// staging only HALF of the trace slice (128 of the 256 GEMM columns) -- the tensor cores of the pair read the two
// halves from both shared memories.  Per CTA and K step the operand fetch drops from 4 KB (A) + 8 KB (B) to 4 + 4 KB:
// the shared-memory data path, which the epilogue's loads and reductions share with the MMA operand fetch and which
// the ncu capture of the one-CTA kernel shows as the co-limiter (profiles/r2_matched_filter_tc.md), gets a third of
// its MMA load back.  Roles per CTA as in mf_tc_kernel; the MMA is issued by the even CTA only, the odd CTA's warp 1
// relays "my A image / trace half has landed" to the leader's barriers; completion of the MMAs is multicast to the
// barriers of both CTAs.
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
// Remote arrive with the default semantics (release at CTA scope), as CUTLASS's ClusterBarrier does: what crosses the
// pair here is "an operand landed in shared memory" / "an accumulator was drained from tensor memory", consumed by
// the tensor core's async proxy -- no global-memory ordering is implied.  (A .release.cluster arrive compiles to
// MEMBAR.ALL.GPU, which made every epilogue warp wait for its own reductions to drain: measured 8 % slower.)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tc_commit2(uint64_t *bar) {  // arrives on the barrier at this offset in BOTH CTAs
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
            smem_u32(bar)),
        "h"((uint16_t)3)
        : "memory");
}
__device__ __forceinline__ void tc_mma2_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

__global__ void __launch_bounds__(TC_THREADS, 1) mf_tc_pair_kernel(const TcParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *a_s = smem;
    uint8_t *b_s = smem + p.a_bytes;
    float *inv_s = reinterpret_cast<float *>(b_s + TC_NB * p.b_stage_bytes);  // [2 halves][TC_NI][TC_INV_SLOT]
    uint64_t *bars = reinterpret_cast<uint64_t *>(inv_s + 2 * TC_NI * TC_INV_SLOT);
    uint64_t *a_full = bars + 0, *a_empty = bars + 1;
    uint64_t *acc_full = bars + 2, *acc_empty = bars + 4;               // [2]
    uint64_t *b_full = bars + 6, *b_empty = bars + 6 + TC_NB;           // [TC_NB]
    uint64_t *inv_full = bars + 6 + 2 * TC_NB;                          // [2][TC_NI]
    uint64_t *inv_empty = inv_full + 2 * TC_NI;                         // [2][TC_NI]
    uint64_t *a_peer_full = inv_empty + 2 * TC_NI;                      // leader only: the odd CTA's A image landed
    uint64_t *b_peer_full = a_peer_full + 1;                            // [TC_NB] leader only: its trace half landed
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(b_peer_full + TC_NB);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();           // 0 = leader (issues the MMAs), 1 = its pair
    const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

    if (threadIdx.x == 0) {
        mbar_init(a_full, 1);
        mbar_init(a_empty, 1);
        mbar_init(a_peer_full, 1);
        for (int i = 0; i < 2; i++) {
            mbar_init(acc_full + i, 1);
            mbar_init(acc_empty + i, 2 * TC_EPI_WARPS);  // the epilogue warps of BOTH CTAs (used on the leader)
        }
        for (int i = 0; i < TC_NB; i++) {
            mbar_init(b_full + i, 1);
            mbar_init(b_empty + i, 1);
            mbar_init(b_peer_full + i, 1);
        }
        for (int i = 0; i < 2 * TC_NI; i++) {
            mbar_init(inv_full + i, 1);
            mbar_init(inv_empty + i, 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(TC_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // both CTAs' barriers are initialised before anyone arrives remotely
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int MTP = p.MT >> 1;  // template tile PAIRS (MT is even here)
    const int n_items = p.n_seg * MTP * p.SC * p.n_kc;
    const uint32_t b_bytes = (uint32_t)p.b_slice_bytes;  // (8*127 + Kc) * 2: this CTA's 128 columns

    // item -> (kc, ch, tile pair, segment), same order as the one-CTA kernel
    auto decode = [&](int it, int &kc, int &ch, int &mt, int &seg) {
        kc = it % p.n_kc;
        it /= p.n_kc;
        ch = it % p.SC;
        it /= p.SC;
        mt = 2 * (it % MTP) + (int)rank;
        seg = it / MTP;
    };
    auto pair_active = [&](int ch, int mt) {
        return p.item_active[ch * p.MT + (mt & ~1)] | p.item_active[ch * p.MT + (mt | 1)];
    };

    if (warp == 0) {
        // ===================== producer: this CTA's A image and its half of every trace slice ================
        if (lane == 0) {
            const uint64_t pol_a = l2_policy_evict_first(), pol = l2_policy_evict_normal();
            uint32_t n_item = 0, n_tile = 0;
            for (int it = cluster_id; it < n_items; it += n_clusters) {
                int kc, ch, mt, seg;
                decode(it, kc, ch, mt, seg);
                if (!pair_active(ch, mt)) continue;
                mbar_wait(a_empty, (n_item & 1) ^ 1);
                mbar_expect_tx(a_full, (uint32_t)p.a_bytes);
                const uint8_t *src = p.aimg + ((size_t)(ch * p.MT + mt) * p.n_kc + kc) * p.a_bytes;
                for (int off = 0; off < p.a_bytes; off += 16384) {
                    const int n = (p.a_bytes - off < 16384) ? (p.a_bytes - off) : 16384;
                    bulk_g2s(a_s + off, src + off, (uint32_t)n, a_full, pol_a);
                }
                n_item++;
                const int t_begin = seg * p.seg_tiles;
                const int t_end = (t_begin + p.seg_tiles < p.n_tiles) ? t_begin + p.seg_tiles : p.n_tiles;
                for (int tile = t_begin; tile < t_end; tile++, n_tile++) {
                    const int s = n_tile % TC_NB;
                    mbar_wait(b_empty + s, ((n_tile / TC_NB) & 1) ^ 1);
                    const size_t d_off = (size_t)ch * p.dh_pitch + ((size_t)tile * TC_NCOL + 128 * rank) * 8 + (size_t)kc * p.Kc;
                    mbar_expect_tx(b_full + s, p.split ? 2 * b_bytes : b_bytes);
                    bulk_g2s(b_s + s * p.b_stage_bytes, p.dh + d_off, b_bytes, b_full + s, pol);
                    if (p.split) bulk_g2s(b_s + s * p.b_stage_bytes + b_bytes, p.dh_lo + d_off, b_bytes, b_full + s, pol);
                }
            }
        }
    } else if (warp == 1 && rank == 1) {
        // ===================== relay (odd CTA): tell the leader when this CTA's operands have landed =========
        if (lane == 0) {
            const uint32_t a_peer = mapa_u32(smem_u32(a_peer_full), 0);
            const uint32_t b_peer = mapa_u32(smem_u32(b_peer_full), 0);
            uint32_t n_item = 0, n_tile = 0;
            for (int it = cluster_id; it < n_items; it += n_clusters) {
                int kc, ch, mt, seg;
                decode(it, kc, ch, mt, seg);
                if (!pair_active(ch, mt)) continue;
                mbar_wait(a_full, n_item & 1);
                mbar_arrive_cluster(a_peer);
                n_item++;
                const int t_begin = seg * p.seg_tiles;
                const int t_end = (t_begin + p.seg_tiles < p.n_tiles) ? t_begin + p.seg_tiles : p.n_tiles;
                for (int k = t_begin; k < t_end; k++, n_tile++) {
                    const int s = n_tile % TC_NB;
                    mbar_wait(b_full + s, (n_tile / TC_NB) & 1);
                    mbar_arrive_cluster(b_peer + 8u * (uint32_t)s);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (one thread of the leader CTA, for the pair) ======================
        if (lane == 0) {
            const uint32_t idesc = tc_idesc(2 * TC_ROWS, TC_NCOL);
            const uint32_t a_addr = smem_u32(a_s), b_addr = smem_u32(b_s);
            const int n_k = p.Kc / 16;
            uint32_t n_item = 0, n_tile = 0;
            for (int it = cluster_id; it < n_items; it += n_clusters) {
                int kc, ch, mt, seg;
                decode(it, kc, ch, mt, seg);
                if (!pair_active(ch, mt)) continue;
                mbar_wait(a_full, n_item & 1);
                mbar_wait(a_peer_full, n_item & 1);
                n_item++;
                const int t_begin = seg * p.seg_tiles;
                const int t_end = (t_begin + p.seg_tiles < p.n_tiles) ? t_begin + p.seg_tiles : p.n_tiles;
                for (int k = t_begin; k < t_end; k++, n_tile++) {
                    const int s = n_tile % TC_NB;
                    const int sa = n_tile & 1;
                    mbar_wait(b_full + s, (n_tile / TC_NB) & 1);
                    mbar_wait(b_peer_full + s, (n_tile / TC_NB) & 1);
                    mbar_wait(acc_empty + sa, ((n_tile >> 1) & 1) ^ 1);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)(sa * TC_NCOL);
                    for (int kk = 0; kk < n_k; kk++) {
                        const uint64_t da = tc_desc(a_addr + (uint32_t)kk * 2 * (TC_ROWS * 16), TC_ROWS * 16, 128);
                        const uint32_t b_at = b_addr + (uint32_t)(s * p.b_stage_bytes) + (uint32_t)kk * 32;
                        const uint64_t db = tc_desc(b_at, 16, 128);
                        tc_mma2_f16(d_tmem, da, db, idesc, kk > 0 ? 1u : 0u);
                        if (p.split) {
                            const uint64_t db_lo = tc_desc(b_at + b_bytes, 16, 128);
                            const uint64_t da_lo = tc_desc(a_addr + (uint32_t)p.a_img_bytes + (uint32_t)kk * 2 * (TC_ROWS * 16),
                                                           TC_ROWS * 16, 128);
                            tc_mma2_f16(d_tmem, da, db_lo, idesc, 1u);
                            tc_mma2_f16(d_tmem, da_lo, db, idesc, 1u);
                        }
                    }
                    tc_commit2(b_empty + s);    // both producers may overwrite their trace halves
                    tc_commit2(acc_full + sa);  // both epilogues may read their accumulators
                }
                tc_commit2(a_empty);
            }
        }
    } else if (warp == TC_EPI_WARP0 + TC_EPI_WARPS) {
        // ===================== rsqrt(energy) producer (per CTA, as in the one-CTA kernel) ===================
        if (lane == 0) {
            const uint64_t pol = l2_policy_evict_normal();
            uint32_t n_chunk = 0;
            for (int it = cluster_id; it < n_items; it += n_clusters) {
                int kc, ch, mt, seg;
                decode(it, kc, ch, mt, seg);
                if (!pair_active(ch, mt)) continue;
                const int t_begin = seg * p.seg_tiles;
                const int t_end = (t_begin + p.seg_tiles < p.n_tiles) ? t_begin + p.seg_tiles : p.n_tiles;
                const float *ci = p.colinv + (size_t)ch * p.ci_pitch;
                for (int tile = t_begin; tile < t_end; tile++) {
                    for (int chunk = 0; chunk < 4; chunk++, n_chunk++) {
                        const int slot = n_chunk % TC_NI;
                        const uint32_t par = ((n_chunk / TC_NI) & 1) ^ 1;
#pragma unroll
                        for (int h = 0; h < 2; h++) {
                            uint64_t *full = inv_full + h * TC_NI + slot;
                            mbar_wait(inv_empty + h * TC_NI + slot, par);
                            mbar_expect_tx(full, TC_INV_SLOT * 4);
                            bulk_g2s(inv_s + (h * TC_NI + slot) * TC_INV_SLOT,
                                     ci + 8 * ((size_t)tile * TC_NCOL + h * 128 + chunk * 32), TC_INV_SLOT * 4, full, pol);
                        }
                    }
                }
            }
        }
    } else {
        // ===================== epilogue (shuffle-free form), this CTA's 128 rows =============================
        const int quarter = warp & 3;
        const int half = ((warp - TC_EPI_WARP0) >> 2) & 1;
        const int sub = (warp - TC_EPI_WARP0) >> 3;
        const uint64_t red_policy = l2_policy_evict_last();
        const int pp = (lane >> 2) & 3;
        const uint32_t acc_empty_leader = mapa_u32(smem_u32(acc_empty), 0);
        uint32_t n_tile = 0;
        for (int it = cluster_id; it < n_items; it += n_clusters) {
            int kc, ch, mt, seg;
            decode(it, kc, ch, mt, seg);
            if (!pair_active(ch, mt)) continue;
            TcPair pr[2];
#pragma unroll
            for (int hh = 0; hh < 2; hh++) {
                const int row0 = quarter * 32 + 16 * hh + 4 * (lane >> 4) + pp;
                const size_t mrow = (size_t)(ch * p.MT + mt) * TC_ROWS + row0;
                const int t = p.row_t[mrow];
                TcPair &r = pr[hh];
                r.scale = p.row_scale[mrow];
                r.ithr = p.row_ithr[mrow];
                r.q = p.row_q[mrow];
                r.ioff = 8 * (lane & 3) + p.row_e[mrow];
                r.ioff_out = 8 * (lane & 3) + 2 * pp;
                r.lo = 1;
                r.hi = 0;
                if (t >= 0 && r.scale != 0.f) { r.lo = p.lo[t]; r.hi = p.hi[t]; }
                r.obase = p.cc + (size_t)(t >= 0 ? t : 0) * p.cc_pitch + r.ioff_out;
                r.jj_full_lo = r.jj_any_lo = 1;
                r.jj_full_hi = r.jj_any_hi = 0;
                if (r.hi >= r.lo) {
                    r.jj_full_lo = ceil_div8(r.lo - 2 * pp) + r.q;
                    r.jj_full_hi = floor_div8(r.hi - 2 * pp - 1) + r.q;
                    r.jj_any_lo = ceil_div8(r.lo - 2 * pp - 1) + r.q;
                    r.jj_any_hi = floor_div8(r.hi - 2 * pp) + r.q;
                }
            }
            const int t_begin = seg * p.seg_tiles;
            const int t_end = (t_begin + p.seg_tiles < p.n_tiles) ? t_begin + p.seg_tiles : p.n_tiles;
            for (int tile = t_begin; tile < t_end; tile++, n_tile++) {
                const int s = n_tile & 1;
                mbar_wait(acc_full + s, (n_tile >> 1) & 1);
                tc_fence_after();
                const int jj0 = tile * TC_NCOL + half * 128;
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(s * TC_NCOL + half * 128);
                int mode[2];
#pragma unroll
                for (int hh = 0; hh < 2; hh++) {
                    const bool full = (pr[hh].hi < pr[hh].lo) || (pr[hh].jj_full_lo <= jj0 && pr[hh].jj_full_hi >= jj0 + 127);
                    const bool any = pr[hh].jj_any_lo <= jj0 + 127 && pr[hh].jj_any_hi >= jj0;
                    mode[hh] = !__any_sync(0xffffffffu, any) ? 0 : (__all_sync(0xffffffffu, full) ? 2 : 1);
                }
                if (p.debug & 1) mode[0] = mode[1] = 0;
#pragma unroll 1
                for (int chunk = sub; chunk < 4; chunk += TC_EPI_SUBS) {
                    const uint32_t n_chunk = 4 * n_tile + (uint32_t)chunk;
                    const int slot = n_chunk % TC_NI;
                    uint64_t *full = inv_full + half * TC_NI + slot;
                    uint32_t r0[16], r1[16];
                    tc_ld16x128_x8(taddr + (uint32_t)(chunk * 32), r0);
                    tc_ld16x128_x8(taddr + (16u << 16) + (uint32_t)(chunk * 32), r1);
                    mbar_wait(full, (n_chunk / TC_NI) & 1);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    const float *inv = inv_s + (half * TC_NI + slot) * TC_INV_SLOT;
                    if (mode[0]) tc_epi_half2(r0, inv, pr[0], 8ll * (jj0 + chunk * 32 - pr[0].q), mode[0], red_policy);
                    if (mode[1]) tc_epi_half2(r1, inv, pr[1], 8ll * (jj0 + chunk * 32 - pr[1].q), mode[1], red_policy);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(inv_empty + half * TC_NI + slot);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(acc_empty_leader + 8u * (uint32_t)s);  // the leader's barrier, from both CTAs
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // nobody frees tensor memory or exits while the pair still uses it
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS)
                     : "memory");
    }
}

// ---------------------------------------------------------------- pre-pass kernels
// per-channel max |d| (bit pattern of a non-negative float orders like an unsigned int)
__global__ void tc_absmax_kernel(const float *__restrict__ data, size_t N, unsigned *__restrict__ chmax) {
    const size_t ch = blockIdx.y;
    const float *d = data + ch * N;
    float m = 0.f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (size_t)gridDim.x * blockDim.x) {
        const float a = fabsf(d[i]);
        if (a < INFINITY) m = fmaxf(m, a);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(chmax + ch, __float_as_uint(m));
}

// power-of-two scale that maps max|x| into [2^13, 2^14): exact in fp32, keeps fp16 in normal range
__device__ __forceinline__ float pow2_scale(float amax) {
    if (!(amax > 0.f)) return 1.f;
    int ex;
    frexpf(amax, &ex);  // amax = f * 2^ex, f in [0.5, 1)
    int sh = 14 - ex;
    sh = sh > 100 ? 100 : (sh < -100 ? -100 : sh);
    return ldexpf(1.f, sh);
}

// traces -> scaled fp16 (hi, and the residual lo in split mode), zero padded to dh_pitch; also
// publishes the per-channel scale
__global__ void tc_data_convert_kernel(const float *__restrict__ data, size_t N, size_t dh_pitch,
                                       const unsigned *__restrict__ chmax, __half *__restrict__ dh,
                                       __half *__restrict__ dh_lo, float *__restrict__ ch_scale) {
    const size_t ch = blockIdx.y;
    const float sB = pow2_scale(__uint_as_float(chmax[ch]));
    if (blockIdx.x == 0 && threadIdx.x == 0) ch_scale[ch] = sB;
    const float *d = data + ch * N;
    __half *o = dh + ch * dh_pitch;
    __half *ol = dh_lo ? dh_lo + ch * dh_pitch : nullptr;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < dh_pitch; i += (size_t)gridDim.x * blockDim.x) {
        const float v = (i >= TC_FRONT && i - TC_FRONT < N) ? d[i - TC_FRONT] * sB : 0.f;
        const __half h = __float2half_rn(v);
        o[i] = h;
        if (ol) ol[i] = __float2half_rn(v - __half2float(h));
    }
}

// Tiles are formed per channel: slot k of channel ch is the k-th template (in index order) that weights ch.
// One CTA per channel, ballot prefix sums over the templates in chunks of 256.
__global__ void __launch_bounds__(256)
    tc_slot_kernel(const float *__restrict__ weights, const float *__restrict__ sst, const int *__restrict__ lo,
                   const int *__restrict__ hi, int T, int SC, int n_slots, int *__restrict__ slot_t) {
    __shared__ int warp_tot[8];
    __shared__ int base_s;
    const int ch = blockIdx.x, tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    int *out = slot_t + (size_t)ch * n_slots;
    if (tid == 0) base_s = 0;
    __syncthreads();
    for (int t0 = 0; t0 < T; t0 += 256) {
        const int t = t0 + tid;
        bool used = false;
        if (t < T) {
            const float w = weights[(size_t)t * SC + ch], ss = sst[(size_t)t * SC + ch];
            used = (w != 0.f) && (hi[t] >= lo[t]) && (ss > 0.f) && (ss < INFINITY);
        }
        const unsigned m = __ballot_sync(0xffffffffu, used);
        if (lane == 0) warp_tot[wid] = __popc(m);
        __syncthreads();
        int before = base_s;
        for (int k = 0; k < wid; k++) before += warp_tot[k];
        if (used) out[before + __popc(m & ((1u << lane) - 1u))] = t;
        __syncthreads();
        if (tid == 0) {
            int tot = 0;
            for (int k = 0; k < 8; k++) tot += warp_tot[k];
            base_s += tot;
        }
        __syncthreads();
    }
    for (int k = base_s + tid; k < n_slots; k += 256) out[k] = -1;
}

// One CTA per (template tile, channel), one thread per A row: builds the resident A images (one per
// K chunk; hi then lo in split mode) in UMMA core-matrix order (K group c of 8 elements at
// c*2048 + row*16 bytes) and the per-row epilogue constants.
__global__ void __launch_bounds__(TC_ROWS)
    tc_template_prep_kernel(const float *__restrict__ templates, const int *__restrict__ moveouts,
                            const float *__restrict__ weights, const float *__restrict__ sst,
                            const int *__restrict__ lo, const int *__restrict__ hi,
                            const float *__restrict__ ch_scale, int T, int SC, int L, int n_kc, int Kc, int split,
                            int a_bytes, int a_img_bytes, const int *__restrict__ slot_t,
                            uint8_t *__restrict__ aimg, int *__restrict__ row_q, int *__restrict__ row_e,
                            float *__restrict__ row_scale, float *__restrict__ row_ithr, int *__restrict__ row_t,
                            int *__restrict__ item_active, int debug, int epi) {
    const int mt = blockIdx.x, ch = blockIdx.y, MT = gridDim.x;
    const int row = threadIdx.x;
    int tl, ir;
    tc_row_decode(row, epi, tl, ir);
    const size_t item = (size_t)ch * MT + mt;
    const size_t mrow = item * TC_ROWS + row;
    const int t = slot_t[item * TC_TPL + tl];  // -1: empty slot
    row_t[mrow] = t;

    bool used = false;
    float w = 0.f, ss = 0.f;
    int m = 0;
    if (t >= 0) {
        w = weights[(size_t)t * SC + ch];
        ss = sst[(size_t)t * SC + ch];
        m = moveouts[(size_t)t * SC + ch];
        used = (w != 0.f) && (hi[t] >= lo[t]) && (ss > 0.f) && (ss < INFINITY);
    }
    const int any = __syncthreads_or(used ? 1 : 0);
    if (row == 0) item_active[item] = any;

    // window start of output i = 8*ii + ir is j = i + m = 8*(ii + q) + ir + (m - 8q): the whole-column
    // part q = floor(m / 8) is the same for the template's 8 rows (so the 8 lanes of a column add
    // into ONE aligned 32-byte sector); the rest, e = ir + m mod 8 in [0, 15), is a delay of the row
    const int q0 = (m >= 0) ? (m >> 3) : -((-m + 7) >> 3);
    const int e = ir + (m - 8 * q0);
    int q = q0 + TC_FRONT / 8;  // staged traces start TC_FRONT samples early
    if (debug & 16) q &= ~3;    // triage only (wrong results): every 128-byte reduction run is line-aligned
    const float *tp = templates + ((size_t)(t >= 0 ? t : 0) * SC + ch) * (size_t)L;
    float sA = 1.f;
    if (used) {
        float amax = 0.f;
        for (int k = 0; k < L; k++) amax = fmaxf(amax, fabsf(tp[k]));
        sA = pow2_scale(amax);
    }
    row_q[mrow] = q;
    row_e[mrow] = e;
    // CC = acc / (sA * sB) / sqrt(sst * E); guard sst * E > 1e-6  <=>  rsqrt(E) < sqrt(sst) * 1e3
    row_scale[mrow] = used ? w / (sA * ch_scale[ch] * sqrtf(ss)) : 0.f;
    row_ithr[mrow] = used ? sqrtf(ss) * 1000.f : 0.f;
    if (!any) return;
    for (int kc = 0; kc < n_kc; kc++) {
        uint4 *img = reinterpret_cast<uint4 *>(aimg + (item * n_kc + kc) * (size_t)a_bytes);
        uint4 *img_lo = reinterpret_cast<uint4 *>(aimg + (item * n_kc + kc) * (size_t)a_bytes + a_img_bytes);
        for (int c = 0; c < Kc / 8; c++) {
            __half h[8], l[8];
#pragma unroll
            for (int x = 0; x < 8; x++) {
                const int k = kc * Kc + c * 8 + x - e;
                const float v = (used && k >= 0 && k < L) ? tp[k] * sA : 0.f;
                h[x] = __float2half_rn(v);
                l[x] = __float2half_rn(v - __half2float(h[x]));
            }
            img[(size_t)c * TC_ROWS + row] = *reinterpret_cast<uint4 *>(h);
            if (split) img_lo[(size_t)c * TC_ROWS + row] = *reinterpret_cast<uint4 *>(l);
        }
    }
}

// ---------------------------------------------------------------- host side
// Tile geometry: depends on (L, N) and on the arithmetic (single pass / split), not on T.
struct TcGeom {
    int n_kc, Kc, a_bytes, a_img_bytes, b_slice_bytes, b_stage_bytes, n_tiles;
    size_t smem_bytes, dh_pitch, ci_pitch;
};

// BPMF_TC_PAIR=1 selects the CTA-pair kernel (tcgen05 cta_group::2).  Measured on this pool (profiles/
// r2_matched_filter_tc.md): it needs 25-30 % fewer SM cycles per pass, but under the 1 kW power cap the clock settles
// that much lower (1.05-1.10 GHz against 1.40-1.45 GHz) and the pass takes the same 310-320 ms, with about twice the
// DRAM traffic -- so the one-CTA kernel stays the default and the pair kernel is kept for parts that are not
// power-capped at this point.
static int tc_pair_mode() {
    const char *e = getenv("BPMF_TC_PAIR");
    return (e && atoi(e) == 1) ? 1 : 0;
}

static TcGeom tc_geom(const MfShape &sh, int split, int pair = 0) {
    TcGeom g;
    const int Kp = (int)align_up(sh.L + 14, 16);  // template length + largest row delay
    const int kc_max = split ? TC_KC_MAX_SPLIT : TC_KC_MAX;
    g.n_kc = (Kp + kc_max - 1) / kc_max;
    g.Kc = (int)align_up((size_t)(Kp + g.n_kc - 1) / g.n_kc, 16);
    g.a_img_bytes = TC_ROWS * g.Kc * 2;
    g.a_bytes = g.a_img_bytes * (split ? 2 : 1);
    g.b_slice_bytes = (8 * ((pair ? TC_NCOL / 2 : TC_NCOL) - 1) + g.Kc) * 2;  // a CTA of a pair stages 128 of the 256 columns
    g.b_stage_bytes = (int)align_up((size_t)g.b_slice_bytes * (split ? 2 : 1), 128);
    g.smem_bytes = (size_t)g.a_bytes + TC_NB * (size_t)g.b_stage_bytes + 2 * TC_NI * TC_INV_SLOT * 4 + 320;
    const size_t n_cols = (sh.N - sh.L + TC_FRONT) / 8 + 1;  // GEMM columns that can hold a valid window start
    g.n_tiles = (int)((n_cols + TC_NCOL - 1) / TC_NCOL);
    // last trace slice read ends at (n_tiles-1)*2048 + (n_kc-1)*Kc + 2040 + Kc; the bound below holds for both
    // arithmetics so that one staged copy of the traces serves the single pass and the split
    g.dh_pitch = align_up((size_t)g.n_tiles * (8 * TC_NCOL) + 2 * (size_t)Kp + 128, 64);
    if (g.dh_pitch < align_up(sh.N + TC_FRONT, 64)) g.dh_pitch = align_up(sh.N + TC_FRONT, 64);
    g.ci_pitch = (size_t)g.n_tiles * (8 * TC_NCOL) + 64;
    return g;
}

// trace-side staging: depends only on the traces (and L) -- prepared once, shared by every template batch
struct TcTraceWs {
    __half *dh, *dh_lo;
    float *colinv, *ch_scale;
    unsigned *chmax;
};
static TcTraceWs tc_carve_traces(Workspace &ws, const MfShape &sh, const TcGeom &g, bool want_lo) {
    TcTraceWs w;
    const size_t SC = sh.SC();
    w.ch_scale = ws.take<float>(SC);
    w.chmax = ws.take<unsigned>(SC);
    w.colinv = ws.take<float>(SC * g.ci_pitch);
    w.dh = ws.take<__half>(SC * g.dh_pitch);
    w.dh_lo = want_lo ? ws.take<__half>(SC * g.dh_pitch) : nullptr;
    return w;
}

// template-side staging of one batch of T templates
struct TcTplWs {
    uint8_t *aimg;
    int *row_q, *row_e, *row_t, *slot_t, *item_active;
    float *row_scale, *row_ithr;
    int MT;
};
static TcTplWs tc_carve_templates(Workspace &ws, const MfShape &sh, const TcGeom &g) {
    TcTplWs w;
    const size_t SC = sh.SC();
    w.MT = (int)((sh.T + TC_TPL - 1) / TC_TPL);
    w.MT += w.MT & 1;  // an even number of template tiles: the pair kernel takes them two by two
    w.aimg = ws.take<uint8_t>(SC * (size_t)w.MT * g.n_kc * g.a_bytes);
    w.row_q = ws.take<int>(SC * (size_t)w.MT * TC_ROWS);
    w.row_e = ws.take<int>(SC * (size_t)w.MT * TC_ROWS);
    w.row_t = ws.take<int>(SC * (size_t)w.MT * TC_ROWS);
    w.slot_t = ws.take<int>(SC * (size_t)w.MT * TC_TPL);
    w.row_scale = ws.take<float>(SC * (size_t)w.MT * TC_ROWS);
    w.row_ithr = ws.take<float>(SC * (size_t)w.MT * TC_ROWS);
    w.item_active = ws.take<int>(SC * (size_t)w.MT);
    return w;
}

static int tc_debug_flags() {  // BPMF_TC_DEBUG: performance triage switches (see tc_epi_chunk), 0 in production
    const char *dbg = getenv("BPMF_TC_DEBUG");
    return dbg ? atoi(dbg) : 0;
}

static int tc_epilogue_variant() {  // BPMF_TC_EPI=0 selects the shuffle-transpose epilogue (A/B measurements)
    const char *e = getenv("BPMF_TC_EPI");
    return (e && atoi(e) == 0) ? 0 : 1;
}

bool mf_tc_supported(size_t L, size_t step, int normalize, int network_sum) {
    return step == 1 && normalize == 0 && network_sum != 0 && L >= 8 && L <= 65536;
}

size_t mf_tc_trace_ws_bytes(const MfShape &sh, bool want_lo) {
    Workspace ws(nullptr, 0);
    tc_carve_traces(ws, sh, tc_geom(sh, 0), want_lo);
    return ws.used;
}

size_t mf_tc_template_ws_bytes(const MfShape &sh, int split) {
    Workspace ws(nullptr, 0);
    tc_carve_templates(ws, sh, tc_geom(sh, split));
    return ws.used;
}

// traces: per-channel scale, fp16 copy (hi [+ lo]), rsqrt(window energy)
int mf_tc_prepare_traces(const MfShape &sh, bool want_lo, const float *d_data, Workspace &trace_ws,
                         cudaStream_t st) {
    BPMF_REQUIRE(mf_tc_supported(sh.L, sh.step, sh.normalize, sh.network_sum),
                 "matched filter: shape not supported by the tensor path");
    const TcGeom g = tc_geom(sh, 0);
    TcTraceWs w = tc_carve_traces(trace_ws, sh, g, want_lo);
    BPMF_REQUIRE(trace_ws.ok(), "matched filter: trace workspace too small for the tensor path");
    const size_t SC = sh.SC();
    BPMF_REQUIRE(SC <= 65535, "matched filter: too many channels");
    BPMF_CUDA(cudaMemsetAsync(w.chmax, 0, sizeof(unsigned) * SC, st));
    tc_absmax_kernel<<<dim3(64, (unsigned)SC), 256, 0, st>>>(d_data, sh.N, w.chmax);
    BPMF_LAUNCH_CHECK();
    tc_data_convert_kernel<<<dim3(128, (unsigned)SC), 256, 0, st>>>(d_data, sh.N, g.dh_pitch, w.chmax, w.dh,
                                                                    w.dh_lo, w.ch_scale);
    BPMF_LAUNCH_CHECK();
    BPMF_CUDA(cudaMemsetAsync(w.colinv, 0, sizeof(float) * SC * g.ci_pitch, st));
    return mf_energy(sh, d_data, w.colinv + TC_FRONT, nullptr, g.ci_pitch, true, false, st);
}

int mf_tc_run(const MfShape &sh, int split, const float *d_templates, const int *d_moveouts,
              const float *d_weights, MfMeta meta, Workspace &trace_ws, bool traces_have_lo, Workspace &tpl_ws,
              float *d_cc, size_t cc_pitch, bool zero_cc, cudaStream_t st) {
    BPMF_REQUIRE(mf_tc_supported(sh.L, sh.step, sh.normalize, sh.network_sum),
                 "matched filter: shape not supported by the tensor path");
    BPMF_REQUIRE(!split || traces_have_lo, "matched filter: traces were prepared without the lo part");
    const int epi = tc_epilogue_variant();
    const int pair = (epi == 1) ? tc_pair_mode() : 0;
    const TcGeom g = tc_geom(sh, split, pair);
    TcTraceWs tw = tc_carve_traces(trace_ws, sh, tc_geom(sh, 0), traces_have_lo);
    TcTplWs w = tc_carve_templates(tpl_ws, sh, g);
    BPMF_REQUIRE(trace_ws.ok() && tpl_ws.ok(), "matched filter: workspace too small for the tensor path");
    BPMF_REQUIRE(g.smem_bytes <= TC_SMEM_LIMIT, "matched filter: tensor-path tile exceeds shared memory");
    const size_t SC = sh.SC();

    // templates: per-channel tiles of the templates that weight the channel, resident A images + per-row constants
    tc_slot_kernel<<<(unsigned)SC, 256, 0, st>>>(d_weights, meta.sst, meta.lo, meta.hi, (int)sh.T, (int)SC,
                                                 w.MT * TC_TPL, w.slot_t);
    BPMF_LAUNCH_CHECK();
    tc_template_prep_kernel<<<dim3((unsigned)w.MT, (unsigned)SC), TC_ROWS, 0, st>>>(
        d_templates, d_moveouts, d_weights, meta.sst, meta.lo, meta.hi, tw.ch_scale, (int)sh.T, (int)SC, (int)sh.L,
        g.n_kc, g.Kc, split, g.a_bytes, g.a_img_bytes, w.slot_t, w.aimg, w.row_q, w.row_e, w.row_scale, w.row_ithr,
        w.row_t, w.item_active, tc_debug_flags(), epi);
    BPMF_LAUNCH_CHECK();
    // cc_sum is accumulated with reductions: start from zero (also the value of never-computed outputs); the
    // second pass of a mixed-precision call adds to what the first one left
    if (zero_cc) BPMF_CUDA(cudaMemsetAsync(d_cc, 0, sizeof(float) * sh.T * cc_pitch, st));

    TcParams p;
    p.dh = tw.dh; p.dh_lo = tw.dh_lo; p.dh_pitch = g.dh_pitch;
    p.colinv = tw.colinv; p.ci_pitch = g.ci_pitch;
    p.aimg = w.aimg;
    p.row_q = w.row_q; p.row_e = w.row_e; p.row_scale = w.row_scale; p.row_ithr = w.row_ithr; p.row_t = w.row_t;
    p.item_active = w.item_active;
    p.lo = meta.lo; p.hi = meta.hi;
    p.cc = d_cc; p.cc_pitch = cc_pitch;
    p.T = (int)sh.T; p.SC = (int)SC; p.MT = w.MT; p.n_kc = g.n_kc; p.Kc = g.Kc; p.split = split ? 1 : 0;
    p.a_bytes = g.a_bytes; p.a_img_bytes = g.a_img_bytes;
    p.b_slice_bytes = g.b_slice_bytes; p.b_stage_bytes = g.b_stage_bytes;
    p.n_tiles = g.n_tiles;
    p.seg_tiles = TC_SEG_TILES;
    if (const char *sg = getenv("BPMF_TC_SEG_TILES")) p.seg_tiles = std::max(1, atoi(sg));  // tuning experiments
    p.n_seg = (g.n_tiles + p.seg_tiles - 1) / p.seg_tiles;
    p.debug = tc_debug_flags();
    const long long n_items = (long long)p.n_seg * p.MT * p.SC * p.n_kc;
    BPMF_REQUIRE(n_items < (1ll << 31), "matched filter: too many work items for the tensor path");
    // persistent grid: one CTA per SM, minus the SMs the caller reserved for a concurrent collective
    const int sms = std::max(1, num_sms() - reserved_sms());
    if (pair) {
        const long long n_pair_items = (long long)p.n_seg * (p.MT / 2) * p.SC * p.n_kc;
        BPMF_CUDA(cudaFuncSetAttribute(mf_tc_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem_bytes));
        const int n_clusters = (int)std::min<long long>(n_pair_items, (long long)std::max(1, sms / 2));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(2 * n_clusters));
        cfg.blockDim = dim3(TC_THREADS);
        cfg.dynamicSmemBytes = g.smem_bytes;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        {
            MainKernelTimer timer(st);
            BPMF_CUDA(cudaLaunchKernelEx(&cfg, mf_tc_pair_kernel, p));
        }
        BPMF_LAUNCH_CHECK();
        return BPMF_OK;
    }
    auto kern = epi ? mf_tc_kernel<1> : mf_tc_kernel<0>;
    BPMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem_bytes));
    const int grid = (int)std::min<long long>(n_items, (long long)sms);
    {
        MainKernelTimer timer(st);
        kern<<<grid, TC_THREADS, g.smem_bytes, st>>>(p);
    }
    BPMF_LAUNCH_CHECK();
    return BPMF_OK;
}

}  // namespace bpmf
