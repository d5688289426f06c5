// tools/red_bench.cu -- microbenchmark: throughput of global fp32 reductions into an L2-resident
// working set, for the access patterns the tensor-path epilogue could use.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/red_bench tools/red_bench.cu && /tmp/red_bench
#include <cuda_runtime.h>
#include <stdio.h>

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

int main() {
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
    }
    return 0;
}
