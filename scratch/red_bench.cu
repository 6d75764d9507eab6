// Scratch microbenchmark (not part of the product): throughput of red.global.add.f32 patterns on L2-resident rows.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void red4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red2(float* p, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}
// mode 0: lane = row, 16 consecutive floats per lane as 4 x v4 (the current epilogue pattern), C = 64 -> 4 chunks of 16
// mode 1: 8 lanes per row cover 32 floats (128 B line) per instruction, 4 rows per instruction
// mode 2: lane = row, scalar red (16 per chunk)
// mode 3: 4 lanes per sector, v2 (16x256b-like): row = lane/4, cols 2*(lane%4)
template <int MODE>
__global__ void k(float* out, const int* rows, int n_rows, int iters, int C) {
  const int lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int n_warps = (gridDim.x * blockDim.x) >> 5;
  for (int it = 0; it < iters; ++it) {
    const int base = ((warp + it * n_warps) * 32) % (n_rows - 32);
    if (MODE == 0) {
      float* p = out + (size_t)rows[base + lane] * C;
      for (int c0 = 0; c0 < C; c0 += 16)
        for (int q = 0; q < 16; q += 4) red4(p + c0 + q, 1.f, 1.f, 1.f, 1.f);
    } else if (MODE == 1) {
      for (int r = 0; r < 32; r += 4) {
        float* p = out + (size_t)rows[base + r + (lane >> 3)] * C;
        for (int c0 = 0; c0 < C; c0 += 32) red4(p + c0 + (lane & 7) * 4, 1.f, 1.f, 1.f, 1.f);
      }
    } else if (MODE == 2) {
      float* p = out + (size_t)rows[base + lane] * C;
      for (int c = 0; c < C; ++c) atomicAdd(p + c, 1.f);
    } else {
      for (int r = 0; r < 32; r += 8) {
        float* p = out + (size_t)rows[base + r + (lane >> 2)] * C;
        for (int c0 = 0; c0 < C; c0 += 8) red2(p + c0 + (lane & 3) * 2, 1.f, 1.f);
      }
    }
  }
}
int main() {
  const int n_rows = 20000, C = 64, iters = 200;
  float* out; int* rows;
  cudaMalloc(&out, (size_t)n_rows * C * 4); cudaMemset(out, 0, (size_t)n_rows * C * 4);
  int* h = new int[n_rows];
  for (int i = 0; i < n_rows; ++i) h[i] = (int)((i * 7919ull + (i / 32) * 104729ull) % n_rows);   // scattered, mostly distinct per warp
  cudaMalloc(&rows, n_rows * 4); cudaMemcpy(rows, h, n_rows * 4, cudaMemcpyHostToDevice);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int mode = 0; mode < 4; ++mode) {
    for (int threads : {128, 256}) {
      float best = 1e9;
      for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(a);
        if (mode == 0) k<0><<<148, threads>>>(out, rows, n_rows, iters, C);
        if (mode == 1) k<1><<<148, threads>>>(out, rows, n_rows, iters, C);
        if (mode == 2) k<2><<<148, threads>>>(out, rows, n_rows, iters, C);
        if (mode == 3) k<3><<<148, threads>>>(out, rows, n_rows, iters, C);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); best = ms < best ? ms : best;
      }
      const double bytes = 148.0 * (threads / 32) * iters * 32.0 * C * 4;
      printf("mode %d threads %d: %.3f ms, %.1f GB/s of atomic payload, %.2f us per 128-row x 64-ch item-equivalent per SM\n",
             mode, threads, best, bytes / best / 1e6, best * 1e3 / (iters * (threads / 32) / 4.0));
    }
  }
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
