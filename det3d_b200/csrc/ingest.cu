// nuScenes-style multi-sweep ingest on the device: the raw sweeps of one sample -> one cloud
// [N, n_feat + 1] = (x, y, z, intensity.., time lag), ready for d3b_voxelize.
//
// Reference semantics, det3d/datasets/pipelines/loading.py:
//   read_file   :17-31   raw file = float32 [n, 5], the first n_feat (4) columns are kept
//   remove_close:34-43   sweeps only: drop points with |x| < radius AND |y| < radius (in the sweep's own frame)
//   read_sweep  :46-64   xyz <- (transform_matrix . [x y z 1]^T)[:3] -- float64 matrix times float64-promoted
//                         points, rounded once into the float32 array; time column = time_lag
//   __call__    :98-124  key frame first (no filter, no transform, time 0), then the chosen sweeps in order;
//                         np.concatenate keeps every sweep's point order
// The order-preserving compaction is a chunked scan (1024-point chunks: count -> scan of chunk counts -> assign).
#include "common.cuh"

namespace d3b {
namespace {

constexpr int kIngestChunk = 1024;

struct IngestParams {
  int n_sweeps, raw_stride, n_feat;
  float radius;
  int off[D3B_INGEST_MAX_SWEEPS + 1];          // raw point offsets of the sweeps
  double m[D3B_INGEST_MAX_SWEEPS][12];         // rows 0..2 of the 4x4 transform
  float time_lag[D3B_INGEST_MAX_SWEEPS];
  unsigned char has_transform[D3B_INGEST_MAX_SWEEPS], filter_close[D3B_INGEST_MAX_SWEEPS];
};

__device__ __forceinline__ int sweep_of(const IngestParams& p, int i) {
  int s = 0;
  while (s + 1 < p.n_sweeps && i >= p.off[s + 1]) ++s;
  return s;
}

__device__ __forceinline__ bool keeps(const IngestParams& p, const float* __restrict__ raw, int i, int s) {
  if (!p.filter_close[s]) return true;
  const float x = raw[(size_t)i * p.raw_stride], y = raw[(size_t)i * p.raw_stride + 1];
  return !(fabsf(x) < p.radius && fabsf(y) < p.radius);                                        // :39-41
}

__global__ void __launch_bounds__(256)
ingest_count(const IngestParams p, const float* __restrict__ raw, int* __restrict__ chunk_cnt) {
  __shared__ int warp_sums[8];
  const int n = p.off[p.n_sweeps];
  const int i0 = blockIdx.x * kIngestChunk + threadIdx.x * 4;
  int local = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (i0 + j < n) local += keeps(p, raw, i0 + j, sweep_of(p, i0 + j)) ? 1 : 0;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) local += __shfl_xor_sync(0xffffffffu, local, d);
  if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < 8; ++w) t += warp_sums[w];
    chunk_cnt[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(1024)
ingest_scan(const int* __restrict__ chunk_cnt, int n_chunks, int* __restrict__ chunk_base, int* __restrict__ n_out,
            int out_cap) {
  __shared__ int warp_sums[32];
  __shared__ int running;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) running = 0;
  __syncthreads();
  for (int base = 0; base < n_chunks; base += blockDim.x) {
    const int g = base + threadIdx.x;
    const int v = g < n_chunks ? chunk_cnt[g] : 0;
    int incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += t;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int w = warp_sums[lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, w, d);
        if (lane >= d) w += t;
      }
      warp_sums[lane] = w;
    }
    __syncthreads();
    if (g < n_chunks) chunk_base[g] = running + (warp == 0 ? 0 : warp_sums[warp - 1]) + incl - v;
    const int total = warp_sums[31];
    __syncthreads();
    if (threadIdx.x == 0) running += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) *n_out = running < out_cap ? running : out_cap;
}

__global__ void __launch_bounds__(256)
ingest_emit(const IngestParams p, const float* __restrict__ raw, const int* __restrict__ chunk_base,
            float* __restrict__ out, int out_cap) {
  __shared__ int warp_sums[8];
  const int n = p.off[p.n_sweeps];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i0 = blockIdx.x * kIngestChunk + threadIdx.x * 4;
  unsigned int flags = 0u;
  int sw[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    sw[j] = i0 + j < n ? sweep_of(p, i0 + j) : 0;
    if (i0 + j < n && keeps(p, raw, i0 + j, sw[j])) flags |= 1u << j;
  }
  const int local = __popc(flags);
  int incl = local;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += t;
  }
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  int warp_off = 0;
  for (int w = 0; w < warp; ++w) warp_off += warp_sums[w];
  int r = chunk_base[blockIdx.x] + warp_off + incl - local;
  const int width = p.n_feat + 1;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (!((flags >> j) & 1u)) continue;
    if (r < out_cap) {
      const float* q = raw + (size_t)(i0 + j) * p.raw_stride;
      float* o = out + (size_t)r * width;
      const int s = sw[j];
      if (p.has_transform[s]) {
        // float64 row . [x y z 1], terms added in index order like a plain dot product, one rounding to fp32 (:55-58)
        const double x = (double)q[0], y = (double)q[1], z = (double)q[2];
        const double* m = p.m[s];
#pragma unroll
        for (int c = 0; c < 3; ++c)
          o[c] = (float)(__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(m[4 * c], x), __dmul_rn(m[4 * c + 1], y)),
                                              __dmul_rn(m[4 * c + 2], z)), m[4 * c + 3]));
      } else {
        o[0] = q[0]; o[1] = q[1]; o[2] = q[2];
      }
      for (int c = 3; c < p.n_feat; ++c) o[c] = q[c];
      o[p.n_feat] = p.time_lag[s];
    }
    ++r;
  }
}

}  // namespace
}  // namespace d3b

using namespace d3b;

extern "C" size_t d3b_ingest_workspace_bytes(int32_t n_points_total) {
  if (n_points_total < 0) return 0;
  return align_up(((size_t)n_points_total / kIngestChunk + 2) * 4) * 2;
}

extern "C" int d3b_ingest_sweeps(const float* raw, const int32_t* sweep_offsets, int32_t n_sweeps, int32_t raw_stride,
                                 int32_t n_feat, const double* transforms, const uint8_t* has_transform,
                                 const float* time_lag, const uint8_t* filter_close, float radius, float* out,
                                 int32_t out_cap, int32_t* n_out, void* workspace, size_t workspace_bytes,
                                 void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  D3B_REQUIRE(sweep_offsets && n_out && has_transform && time_lag && filter_close, "d3b_ingest_sweeps: null argument");
  D3B_REQUIRE(n_sweeps >= 1 && n_sweeps <= D3B_INGEST_MAX_SWEEPS, "d3b_ingest_sweeps: %d sweeps outside [1, %d]", n_sweeps,
              D3B_INGEST_MAX_SWEEPS);
  D3B_REQUIRE(n_feat >= 3 && raw_stride >= n_feat && out_cap >= 0, "d3b_ingest_sweeps: bad layout (n_feat %d, stride %d)",
              n_feat, raw_stride);
  IngestParams p;
  p.n_sweeps = n_sweeps; p.raw_stride = raw_stride; p.n_feat = n_feat; p.radius = radius;
  for (int s = 0; s <= n_sweeps; ++s) {
    p.off[s] = sweep_offsets[s];
    D3B_REQUIRE(s == 0 ? p.off[s] == 0 : p.off[s] >= p.off[s - 1], "d3b_ingest_sweeps: sweep_offsets not monotone");
  }
  for (int s = 0; s < n_sweeps; ++s) {
    p.has_transform[s] = has_transform[s] ? 1 : 0;
    p.filter_close[s] = filter_close[s] ? 1 : 0;
    p.time_lag[s] = time_lag[s];
    D3B_REQUIRE(!p.has_transform[s] || transforms, "d3b_ingest_sweeps: transforms missing");
    for (int c = 0; c < 12; ++c) p.m[s][c] = p.has_transform[s] ? transforms[(size_t)s * 16 + c] : 0.0;
  }
  const int n = p.off[n_sweeps];
  if (n == 0) {
    D3B_CUDA(cudaMemsetAsync(n_out, 0, 4, stream));
    return D3B_OK;
  }
  D3B_REQUIRE(raw && out && workspace, "d3b_ingest_sweeps: null buffer");
  const size_t need = d3b_ingest_workspace_bytes(n);
  if (need > workspace_bytes) {
    set_error("d3b_ingest_sweeps: workspace %zu < %zu", workspace_bytes, need);
    return D3B_ERR_WORKSPACE;
  }
  const int n_chunks = div_up(n, kIngestChunk);
  int* chunk_cnt = (int*)workspace;
  int* chunk_base = (int*)((char*)workspace + need / 2);
  ingest_count<<<n_chunks, 256, 0, stream>>>(p, raw, chunk_cnt);
  D3B_LAUNCH_CHECK();
  ingest_scan<<<1, 1024, 0, stream>>>(chunk_cnt, n_chunks, chunk_base, n_out, out_cap);
  D3B_LAUNCH_CHECK();
  ingest_emit<<<n_chunks, 256, 0, stream>>>(p, raw, chunk_base, out, out_cap);
  D3B_LAUNCH_CHECK();
  return D3B_OK;
}
