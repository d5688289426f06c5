// tools/red_bench.cu -- microbenchmark: throughput of global fp32 reductions into an L2-resident
// working set, for the access patterns the tensor-path epilogue could use.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/red_bench tools/red_bench.cu && /tmp/red_bench
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

// pattern A: warp = 4 rows x 8 consecutive floats (scalar red per lane, 4 full sectors per warp op)
__global__ void red_scalar(float* out, size_t pitch, int rows, int cols, int iters) {
    const int lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int r = (warp * 4 + (lane >> 3)) % rows;
    for (int it = 0; it < iters; it++) {
        float* p = out + (size_t)r * pitch + (size_t)((it * 977 + warp / (rows / 4)) % (cols / 8)) * 8 + (lane & 7);
#pragma unroll
        for (int u = 0; u < 32; u++) atomicAdd(p + (u * 8) % (cols - 8), 1.0f);
    }
}
// pattern B: each lane red.v4 (16 B); lane pairs adjacent -> 16 rows x 32 B per warp op
__global__ void red_v4(float* out, size_t pitch, int rows, int cols, int iters) {
    const int lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int r = (warp * 16 + (lane >> 1)) % rows;
    for (int it = 0; it < iters; it++) {
        float* p = out + (size_t)r * pitch + (size_t)((it * 977 + warp / (rows / 16)) % (cols / 8)) * 8 + (lane & 1) * 4;
#pragma unroll
        for (int u = 0; u < 32; u++) {
            float* q = p + (u * 8) % (cols - 8);
            asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(q), "f"(1.f), "f"(1.f), "f"(1.f), "f"(1.f) : "memory");
        }
    }
}
// pattern C: each lane red.v4, one row per lane (32 rows x 16 B per warp op)
__global__ void red_v4_rows(float* out, size_t pitch, int rows, int cols, int iters) {
    const int lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int r = (warp * 32 + lane) % rows;
    for (int it = 0; it < iters; it++) {
        float* p = out + (size_t)r * pitch + (size_t)((it * 977 + warp / (rows / 32 + 1)) % (cols / 4)) * 4;
#pragma unroll
        for (int u = 0; u < 32; u++) {
            float* q = p + (u * 4) % (cols - 4);
            asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(q), "f"(1.f), "f"(1.f), "f"(1.f), "f"(1.f) : "memory");
        }
    }
}
// pattern D: plain st.v4, one row per lane (reference for store bandwidth)
__global__ void st_v4_rows(float* out, size_t pitch, int rows, int cols, int iters) {
    const int lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int r = (warp * 32 + lane) % rows;
    for (int it = 0; it < iters; it++) {
        float* p = out + (size_t)r * pitch + (size_t)((it * 977 + warp / (rows / 32 + 1)) % (cols / 4)) * 4;
#pragma unroll
        for (int u = 0; u < 32; u++) *reinterpret_cast<float4*>(p + (u * 4) % (cols - 4)) = make_float4(1, 1, 1, 1);
    }
}

// pattern E: warp = 1 row x 32 consecutive floats (one full 128-byte line per warp op)
__global__ void red_line(float* out, size_t pitch, int rows, int cols, int iters) {
    const int lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int r = warp % rows;
    for (int it = 0; it < iters; it++) {
        float* p = out + (size_t)r * pitch + (size_t)((it * 977 + warp / rows) % (cols / 32 - 32)) * 32 + lane;
#pragma unroll
        for (int u = 0; u < 32; u++) atomicAdd(p + u * 32, 1.0f);
    }
}
// pattern F: warp = 1 row, 4 sectors alternating over 2 lines (lane group g -> sector 2g), as the
// 16x256b TMEM load shape would give
__global__ void red_alt(float* out, size_t pitch, int rows, int cols, int iters) {
    const int lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int r = warp % rows;
    for (int it = 0; it < iters; it++) {
        float* p = out + (size_t)r * pitch + (size_t)((it * 977 + warp / rows) % (cols / 64 - 32)) * 64 + (lane >> 3) * 16 + (lane & 7);
#pragma unroll
        for (int u = 0; u < 32; u++) atomicAdd(p + (u >> 1) * 64 + (u & 1) * 8, 1.0f);
    }
}
// pattern G: red.v4, 8 lanes per 128-byte line, 4 rows per warp op (512 B per op)
__global__ void red_v4_line(float* out, size_t pitch, int rows, int cols, int iters) {
    const int lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int r = (warp * 4 + (lane >> 3)) % rows;
    for (int it = 0; it < iters; it++) {
        float* p = out + (size_t)r * pitch + (size_t)((it * 977 + warp / (rows / 4)) % (cols / 32 - 32)) * 32 + (lane & 7) * 4;
#pragma unroll
        for (int u = 0; u < 32; u++) {
            float* q = p + u * 32;
            asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(q), "f"(1.f), "f"(1.f), "f"(1.f), "f"(1.f) : "memory");
        }
    }
}

// pattern H: red.v2.f32, 16 lanes per 128-byte run, 2 rows per warp op (256 B per op): the shuffle-free epilogue
__global__ void red_v2_line(float* out, size_t pitch, int rows, int cols, int iters) {
    const int lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int r = (warp * 2 + (lane >> 4)) % rows;
    for (int it = 0; it < iters; it++) {
        float* p = out + (size_t)r * pitch + (size_t)((it * 977 + warp / (rows / 2)) % (cols / 32 - 32)) * 32 + (lane & 15) * 2;
#pragma unroll
        for (int u = 0; u < 32; u++) {
            float* q = p + u * 32;
            asm volatile("red.global.add.v2.f32 [%0], {%1,%2};" ::"l"(q), "f"(1.f), "f"(1.f) : "memory");
        }
    }
}
// pattern I: red.global.add.u64 (three 21-bit fixed-point partial sums per word): 8 lanes x 8 B = 64 B per row,
// 4 rows per warp op (256 B per op)
__global__ void red_u64_rows(float* out, size_t pitch, int rows, int cols, int iters) {
    const int lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int r = (warp * 4 + (lane >> 3)) % rows;
    for (int it = 0; it < iters; it++) {
        unsigned long long* p = reinterpret_cast<unsigned long long*>(out + (size_t)r * pitch) +
                                (size_t)((it * 977 + warp / (rows / 4)) % (cols / 16 - 32)) * 8 + (lane & 7);
#pragma unroll
        for (int u = 0; u < 32; u++) {
            unsigned long long* q = p + u * 8;
            asm volatile("red.global.add.u64 [%0], %1;" ::"l"(q), "l"(0x0000040000200001ull) : "memory");
        }
    }
}
// pattern J: red.global.add.u64, 16 lanes x 8 B = one 128-byte line per row, 2 rows per warp op
__global__ void red_u64_line(float* out, size_t pitch, int rows, int cols, int iters) {
    const int lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int r = (warp * 2 + (lane >> 4)) % rows;
    for (int it = 0; it < iters; it++) {
        unsigned long long* p = reinterpret_cast<unsigned long long*>(out + (size_t)r * pitch) +
                                (size_t)((it * 977 + warp / (rows / 2)) % (cols / 32 - 32)) * 16 + (lane & 15);
#pragma unroll
        for (int u = 0; u < 32; u++) {
            unsigned long long* q = p + u * 16;
            asm volatile("red.global.add.u64 [%0], %1;" ::"l"(q), "l"(0x0000040000200001ull) : "memory");
        }
    }
}

template <typename F>
void run(const char* name, F launch, double elems_per_launch) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    launch(); cudaDeviceSynchronize();
    cudaEventRecord(a);
    for (int i = 0; i < 5; i++) launch();
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); ms /= 5;
    printf("%-14s %8.3f ms  %8.2f Gelem/s  %8.1f GB/s(4B/elem)  err=%s\n", name, ms, elems_per_launch / ms / 1e6,
           elems_per_launch * 4 / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main(int argc, char** argv) {
    const int rows = 512, cols = 16384;  // 32 MB working set: L2 resident
    const size_t pitch = cols;
    float* out; cudaMalloc(&out, sizeof(float) * rows * pitch); cudaMemset(out, 0, sizeof(float) * rows * pitch);
    for (int threads : {256}) {
        const int blocks = 148, iters = 64;
        const double warps = blocks * threads / 32.0;
        printf("--- %d CTAs x %d threads\n", blocks, threads);
        run("red.f32 4x8", [&] { red_scalar<<<blocks, threads>>>(out, pitch, rows, cols, iters); }, warps * iters * 32 * 32);
        run("red.v4 16x8", [&] { red_v4<<<blocks, threads>>>(out, pitch, rows, cols, iters); }, warps * iters * 32 * 128);
        run("red.v4 32x4", [&] { red_v4_rows<<<blocks, threads>>>(out, pitch, rows, cols, iters); }, warps * iters * 32 * 128);
        run("red.f32 1x32", [&] { red_line<<<blocks, threads>>>(out, pitch, rows, cols, iters); }, warps * iters * 32 * 32);
        run("red.f32 alt", [&] { red_alt<<<blocks, threads>>>(out, pitch, rows, cols, iters); }, warps * iters * 32 * 32);
        run("red.v4 4x32", [&] { red_v4_line<<<blocks, threads>>>(out, pitch, rows, cols, iters); }, warps * iters * 32 * 128);
        run("st.v4 32x4", [&] { st_v4_rows<<<blocks, threads>>>(out, pitch, rows, cols, iters); }, warps * iters * 32 * 128);
        run("red.v2 2x32", [&] { red_v2_line<<<blocks, threads>>>(out, pitch, rows, cols, iters); }, warps * iters * 32 * 64);
        // u64 rows: "elements" are counted as 4-byte units moved, so that GB/s is comparable (8 B per lane per op)
        run("red.u64 4x8", [&] { red_u64_rows<<<blocks, threads>>>(out, pitch, rows, cols, iters); }, warps * iters * 32 * 64);
        run("red.u64 2x16", [&] { red_u64_line<<<blocks, threads>>>(out, pitch, rows, cols, iters); }, warps * iters * 32 * 64);
    }
    // energy per byte: run one pattern for about `secs` seconds (argv: pattern name, seconds) while the caller samples power
    if (argc >= 3) {
        const int blocks = 148, threads = 256, iters = 64;
        const double secs = atof(argv[2]);
        const double warps = blocks * threads / 32.0;
        cudaEvent_t a, b;
        cudaEventCreate(&a); cudaEventCreate(&b);
        cudaEventRecord(a);
        double bytes = 0;
        float ms = 0;
        do {
            for (int i = 0; i < 200; i++) {
                if (!strcmp(argv[1], "red.v2")) { red_v2_line<<<blocks, threads>>>(out, pitch, rows, cols, iters); bytes += warps * iters * 32 * 64 * 4; }
                else if (!strcmp(argv[1], "red.v4")) { red_v4_line<<<blocks, threads>>>(out, pitch, rows, cols, iters); bytes += warps * iters * 32 * 128 * 4; }
                else if (!strcmp(argv[1], "st.v4")) { st_v4_rows<<<blocks, threads>>>(out, pitch, rows, cols, iters); bytes += warps * iters * 32 * 128 * 4; }
                else if (!strcmp(argv[1], "red.u64")) { red_u64_rows<<<blocks, threads>>>(out, pitch, rows, cols, iters); bytes += warps * iters * 32 * 64 * 4; }
                else { red_scalar<<<blocks, threads>>>(out, pitch, rows, cols, iters); bytes += warps * iters * 32 * 32 * 4; }
            }
            cudaEventRecord(b); cudaEventSynchronize(b);
            cudaEventElapsedTime(&ms, a, b);
        } while (ms < 1e3 * secs);
        printf("SUSTAINED %s: %.2f s, %.1f GB moved, %.1f GB/s\n", argv[1], ms / 1e3, bytes / 1e9, bytes / 1e6 / ms);
    }
    return 0;
}
