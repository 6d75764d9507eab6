// Probe: shared-memory cost of the 16-byte cp.async row gather for different thread -> (row, chunk) mappings.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// VAR 0: thread (g = lane/8, c = lane%8) -> rows g*16 + q (current kernel)      [instruction touches rows r, r+16, r+32, r+48]
// VAR 1: rows q*4 + g                                                            [instruction touches 4 consecutive rows]
// VAR 2: like 0 but cp.async.ca
// VAR 3: like 0, ld.global.v4 + st.shared.v4 (registers)
// VAR 4: thread (g = lane/16 .. 2 rows per instr, c = lane%16: 8-byte chunks) cp.async.ca 8 bytes
// VAR 5: like 1 but one plane per instruction pair interleaved differently: hi for all rows first, then lo
template <int VAR>
__global__ void __launch_bounds__(256) tput(const __half* table, const int* idx, int n_idx, int slots, long long* cycles) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  const int group = threadIdx.x >> 6, t = threadIdx.x & 63;
  const int wq = t >> 5, lane = t & 31;
  const int g = lane >> 3, c = lane & 7;
  const long long t0 = clock64();
  for (int j = group; j < slots; j += 4) {
    const uint32_t stage = base + group * 32768;
    const int* ib = idx + ((size_t)(blockIdx.x * slots + j) * 128) % n_idx;
    if (VAR == 0 || VAR == 2 || VAR == 3) {
      const int4* idx4 = reinterpret_cast<const int4*>(ib + wq * 64 + g * 16);
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        const int4 iv = idx4[q4];
        const int srcs[4] = {iv.x, iv.y, iv.z, iv.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int row = wq * 64 + g * 16 + q4 * 4 + u;
          const bool live = srcs[u] >= 0;
          const __half* src = table + (live ? (size_t)srcs[u] * 64 + c * 8 : 0);
          const uint32_t dst = stage + row * 128 + ((c ^ (row & 7)) << 4);
          if (VAR == 0) {
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(live ? 16u : 0u) : "memory");
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + 16384), "l"(src), "r"(live ? 16u : 0u) : "memory");
          } else if (VAR == 2) {
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(live ? 16u : 0u) : "memory");
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst + 16384), "l"(src), "r"(live ? 16u : 0u) : "memory");
          } else {
            const uint4 v = live ? __ldg(reinterpret_cast<const uint4*>(src)) : make_uint4(0, 0, 0, 0);
            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(dst + 16384), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
          }
        }
      }
    } else if (VAR == 1 || VAR == 5) {
      int srcs[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) srcs[q] = ib[wq * 64 + q * 4 + g];
      if (VAR == 1) {
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          const int row = wq * 64 + q * 4 + g;
          const bool live = srcs[q] >= 0;
          const __half* src = table + (live ? (size_t)srcs[q] * 64 + c * 8 : 0);
          const uint32_t dst = stage + row * 128 + ((c ^ (row & 7)) << 4);
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(live ? 16u : 0u) : "memory");
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + 16384), "l"(src), "r"(live ? 16u : 0u) : "memory");
        }
      } else {
#pragma unroll
        for (int pl = 0; pl < 2; ++pl)
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            const int row = wq * 64 + q * 4 + g;
            const bool live = srcs[q] >= 0;
            const __half* src = table + (live ? (size_t)srcs[q] * 64 + c * 8 : 0);
            const uint32_t dst = stage + pl * 16384 + row * 128 + ((c ^ (row & 7)) << 4);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(live ? 16u : 0u) : "memory");
          }
      }
    } else if (VAR == 4) {
      const int g2 = lane >> 4, c2 = lane & 15;        // 2 rows per instruction, 8-byte chunks
#pragma unroll 8
      for (int q = 0; q < 32; ++q) {
        const int row = wq * 64 + q * 2 + g2;
        const int s = ib[row];
        const bool live = s >= 0;
        const __half* src = table + (live ? (size_t)s * 64 + c2 * 4 : 0);
        const uint32_t dst = stage + row * 128 + (((c2 >> 1) ^ (row & 7)) << 4) + (c2 & 1) * 8;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(live ? 8u : 0u) : "memory");
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst + 16384), "l"(src), "r"(live ? 8u : 0u) : "memory");
      }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
}

template <int VAR>
void run(const char* name, const __half* table, const int* d_big, int n_idx, int slots, long long* d_cyc) {
  CK(cudaFuncSetAttribute(tput<VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 32768 + 2048));
  for (int rep = 0; rep < 2; ++rep) {
    tput<VAR><<<148, 256, 4 * 32768 + 2048>>>(table, d_big, n_idx, slots, d_cyc);
    CK(cudaDeviceSynchronize());
  }
  long long c[148]; CK(cudaMemcpy(c, d_cyc, sizeof(c), cudaMemcpyDeviceToHost));
  long long m = 0; for (int i = 0; i < 148; ++i) m = c[i] > m ? c[i] : m;
  printf("  %-44s %.0f cycles/slot\n", name, (double)m / slots);
}

int main() {
  const int R = 16384, slots = 2000, n_idx = 1 << 20;
  __half* table; CK(cudaMalloc(&table, (size_t)R * 64 * 2)); CK(cudaMemset(table, 0, (size_t)R * 64 * 2));
  std::vector<int> big(n_idx);
  long long* d_cyc; CK(cudaMalloc(&d_cyc, 148 * 8));
  int* d_big; CK(cudaMalloc(&d_big, n_idx * 4));
  for (int mode = 0; mode < 3; ++mode) {
    srand(7);
    for (int i = 0; i < n_idx; ++i) {
      const int r = mode == 1 ? (i * 3) % R : rand() % R;
      big[i] = (mode == 2 && (rand() % 4) != 0) ? -1 : r;          // mode 2: 75 % of the rows absent (zero fill)
    }
    CK(cudaMemcpy(d_big, big.data(), n_idx * 4, cudaMemcpyHostToDevice));
    printf("%s\n", mode == 0 ? "random rows" : mode == 1 ? "streaming rows" : "random rows, 75 % absent");
    run<0>("v0 rows r,r+16,r+32,r+48 per instr (cg)", table, d_big, n_idx, slots, d_cyc);
    run<1>("v1 four consecutive rows per instr (cg)", table, d_big, n_idx, slots, d_cyc);
    run<5>("v5 consecutive rows, plane by plane (cg)", table, d_big, n_idx, slots, d_cyc);
    run<2>("v2 as v0 with cp.async.ca", table, d_big, n_idx, slots, d_cyc);
    run<3>("v3 ld.global.v4 + st.shared.v4", table, d_big, n_idx, slots, d_cyc);
    run<4>("v4 8-byte cp.async.ca, two rows per instr", table, d_big, n_idx, slots, d_cyc);
  }
  return 0;
}
