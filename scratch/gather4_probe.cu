// Probe: does TMA tile::gather4 deliver what the sparse kernel's A operand needs?
//  (1) layout: 4-row gathers placed at 512-byte steps inside a 128B-swizzled K-major tile == the canonical layout?
//  (2) a negative row index is zero-filled?
//  (3) throughput per SM against the 16-byte cp.async gather.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint32_t b, uint32_t tx) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(tx) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t b, uint32_t ph) {
  asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}" ::"r"(b), "r"(ph) : "memory");
}
__device__ __forceinline__ void gather4(uint32_t dst, const CUtensorMap* tm, int c0, int r0, int r1, int r2, int r3, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
               ::"r"(dst), "l"(tm), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar) : "memory");
}

__global__ void layout_kernel(const __grid_constant__ CUtensorMap tm, const int* idx, uint8_t* out, int row_bytes) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* gen = raw + (base - smem_u32(raw));
  const uint32_t bar = base + 128 * 128;
  for (int i = threadIdx.x; i < 128 * 128 / 4; i += 32) reinterpret_cast<uint32_t*>(gen)[i] = 0xdeadbeefu;
  if (threadIdx.x == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  asm volatile("fence.proxy.async;" ::: "memory");
  __syncwarp();
  const int l = threadIdx.x;
  if (l == 0) mbar_expect(bar, 128 * row_bytes);
  __syncwarp();
  gather4(base + l * 4 * row_bytes, &tm, 0, idx[4 * l], idx[4 * l + 1], idx[4 * l + 2], idx[4 * l + 3], bar);
  mbar_wait(bar, 0);
  for (int i = threadIdx.x; i < 128 * 128; i += 32) out[i] = gen[i];
}

// throughput: every CTA gathers `slots` tiles of 128 rows x 64 f16 x 2 planes (32 KB), 4 stages in flight
__global__ void __launch_bounds__(32) tput_tma(const __grid_constant__ CUtensorMap tm, const int* idx, int n_idx, int slots, long long* cycles) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t bars = base + 4 * 32768;
  if (threadIdx.x == 0) { for (int s = 0; s < 4; ++s) mbar_init(bars + 8 * s, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncwarp();
  const int l = threadIdx.x;
  const long long t0 = clock64();
  for (int j = 0; j < slots + 3; ++j) {
    if (j < slots) {
      const int s = j & 3;
      const int4 iv = *reinterpret_cast<const int4*>(idx + ((size_t)(blockIdx.x * slots + j) * 128 + 4 * l) % n_idx);
      if (l == 0) mbar_expect(bars + 8 * s, 32768);
      __syncwarp();
      gather4(base + s * 32768 + l * 512, &tm, 0, iv.x, iv.y, iv.z, iv.w, bars + 8 * s);
      gather4(base + s * 32768 + 16384 + l * 512, &tm, 0, iv.x, iv.y, iv.z, iv.w, bars + 8 * s);
    }
    if (j >= 3) mbar_wait(bars + 8 * ((j - 3) & 3), ((j - 3) >> 2) & 1);
    __syncwarp();
  }
  if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
}

// the current scheme: 64 threads, 16 rows x one 16-byte chunk each, both planes, cp.async; 4 groups -> 4 slots in flight
__global__ void __launch_bounds__(256) tput_cpasync(const __half* table, const int* idx, int n_idx, int slots, long long* cycles) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  const int group = threadIdx.x >> 6, t = threadIdx.x & 63;
  const int g = t >> 3, c = t & 7;
  const long long t0 = clock64();
  for (int j = group; j < slots; j += 4) {
    const uint32_t stage = base + group * 32768;
    const int4* idx4 = reinterpret_cast<const int4*>(idx + ((size_t)(blockIdx.x * slots + j) * 128 + g * 16) % n_idx);
#pragma unroll
    for (int q4 = 0; q4 < 4; ++q4) {
      const int4 iv = idx4[q4];
      const int srcs[4] = {iv.x, iv.y, iv.z, iv.w};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int row = g * 16 + q4 * 4 + u;
        const bool live = srcs[u] >= 0;
        const __half* src = table + (live ? (size_t)srcs[u] * 64 + c * 8 : 0);
        const uint32_t dst = stage + row * 128 + ((c ^ (row & 7)) << 4);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(live ? 16u : 0u) : "memory");
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + 16384), "l"(src), "r"(live ? 16u : 0u) : "memory");
      }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  EncodeFn encode = (EncodeFn)fn;
  const int R = 16384;
  for (int cin : {64, 32, 16}) {
    std::vector<__half> h((size_t)R * cin);
    for (int r = 0; r < R; ++r) for (int c = 0; c < cin; ++c) h[(size_t)r * cin + c] = __float2half((float)((r * 7 + c) % 2048));
    __half* table; CK(cudaMalloc(&table, h.size() * 2)); CK(cudaMemcpy(table, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
    std::vector<int> idx(128);
    srand(1);
    for (int i = 0; i < 128; ++i) idx[i] = (i % 5 == 3) ? -1 : rand() % R;
    idx[7] = R;          // one past the end: also out of bounds
    int* d_idx; CK(cudaMalloc(&d_idx, 512)); CK(cudaMemcpy(d_idx, idx.data(), 512, cudaMemcpyHostToDevice));
    uint8_t* d_out; CK(cudaMalloc(&d_out, 16384));
    const int row_bytes = cin * 2;
    const CUtensorMapSwizzle sw = cin == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : cin == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
    const int sw_mask = cin == 64 ? 7 : cin == 32 ? 3 : 1;      // chunk ^= (row-in-atom bits)
    for (int box_rows : {1, 4}) {
      CUtensorMap tm;
      cuuint64_t dims[2] = {(cuuint64_t)cin, (cuuint64_t)R};
      cuuint64_t strides[1] = {(cuuint64_t)cin * 2};
      cuuint32_t box[2] = {(cuuint32_t)cin, (cuuint32_t)box_rows};
      cuuint32_t es[2] = {1, 1};
      CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, table, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      printf("cin %d box_rows %d: encode -> %d\n", cin, box_rows, (int)r);
      if (r != CUDA_SUCCESS) continue;
      CK(cudaMemset(d_out, 0xab, 16384));
      CK(cudaFuncSetAttribute(layout_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 20000));
      layout_kernel<<<1, 32, 20000>>>(tm, d_idx, d_out, row_bytes);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("  kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
      std::vector<uint8_t> out(16384);
      CK(cudaMemcpy(out.data(), d_out, 16384, cudaMemcpyDeviceToHost));
      // expected canonical K-major layout: row pitch = row_bytes, 16-byte chunk c of row r at chunk (c ^ f(r))
      int bad = 0, bad_zero = 0;
      const int chunks = row_bytes / 16;
      for (int rr = 0; rr < 128; ++rr) for (int c = 0; c < chunks; ++c) {
        // swizzle: Swizzle<B,4,3> on the byte offset: bits [4, 4+B) ^= bits [7, 7+B)
        const int lin = rr * row_bytes + c * 16;
        const int off = lin ^ (((lin >> 7) & sw_mask) << 4);
        const __half* got = reinterpret_cast<const __half*>(out.data() + off);
        const bool live = idx[rr] >= 0 && idx[rr] < R;
        for (int k = 0; k < 8; ++k) {
          const float want = live ? __half2float(h[(size_t)idx[rr] * cin + c * 8 + k]) : 0.f;
          if (__half2float(got[k]) != want) { ++bad; if (!live) ++bad_zero; if (bad < 4) printf("    row %d chunk %d k %d: got %g want %g\n", rr, c, k, __half2float(got[k]), want); }
        }
      }
      printf("  layout mismatches: %d (in out-of-bounds rows: %d)\n", bad, bad_zero);
      if (cin == 64 && bad == 0) {
        const int slots = 2000, n_idx = 1 << 20;
        std::vector<int> big(n_idx);
        for (int mode = 0; mode < 2; ++mode) {
          for (int i = 0; i < n_idx; ++i) big[i] = mode == 0 ? rand() % R : ((i * 3) % R);   // random / streaming
          int* d_big; CK(cudaMalloc(&d_big, n_idx * 4)); CK(cudaMemcpy(d_big, big.data(), n_idx * 4, cudaMemcpyHostToDevice));
          long long* d_cyc; CK(cudaMalloc(&d_cyc, 148 * 8));
          CK(cudaFuncSetAttribute(tput_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 32768 + 2048));
          CK(cudaFuncSetAttribute(tput_cpasync, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 32768 + 2048));
          for (int rep = 0; rep < 2; ++rep) {
            tput_tma<<<148, 32, 4 * 32768 + 2048>>>(tm, d_big, n_idx, slots, d_cyc);
            CK(cudaDeviceSynchronize());
            long long c1[148]; CK(cudaMemcpy(c1, d_cyc, sizeof(c1), cudaMemcpyDeviceToHost));
            tput_cpasync<<<148, 256, 4 * 32768 + 2048>>>(table, d_big, n_idx, slots, d_cyc);
            CK(cudaDeviceSynchronize());
            long long c2[148]; CK(cudaMemcpy(c2, d_cyc, sizeof(c2), cudaMemcpyDeviceToHost));
            long long m1 = 0, m2 = 0; for (int i = 0; i < 148; ++i) { m1 = c1[i] > m1 ? c1[i] : m1; m2 = c2[i] > m2 ? c2[i] : m2; }
            printf("  %s rows: gather4 %.0f cycles/slot, cp.async %.0f cycles/slot (32 KB per slot, 148 CTAs)\n",
                   mode == 0 ? "random" : "streaming", (double)m1 / slots, (double)m2 / slots);
          }
        }
      }
    }
  }
  return 0;
}
