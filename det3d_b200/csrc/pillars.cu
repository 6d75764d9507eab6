// PillarFeatureNet in one kernel (eval mode, single PFN layer -- the shape every Det3D PointPillars config
// uses): 9-feature decoration, Linear(ndim+5 -> units, no bias), folded BatchNorm1d, ReLU and the max over the
// pillar's point slots.  Semantics follow det3d/models/readers/pillar_encoder.py:115-155 (decoration + padding
// mask) and :33-47 (PFNLayer); the reference materialises [M, P, 9] and [M, P, units] tensors in HBM, here
// they only ever exist in registers.
//
// One warp per pillar.  The pillar's <= P points are staged once in shared memory (coalesced), every lane owns
// units/32 output channels and walks the points with broadcast LDS reads.  Bound: HBM read of the voxel buffer
// (P*ndim*4 B per pillar) -- the arithmetic is 2*P*(ndim+5)*units flop per pillar, two orders below the fp32
// roof at that byte count.
#include "common.cuh"

namespace d3b {
namespace {

constexpr int kPillarWarps = 4;
constexpr int kMaxUnitsPerLane = 4;   // units <= 128
constexpr int kMaxIn = 16;            // ndim + 5 <= 16
constexpr int kPillarMaxBatch = 64;
constexpr int kListSentinel = 0x7f000000;   // voxelize.cu: list slots >= this are empty

// Where a pillar's points come from: the materialised [rows, P, ndim] voxel buffer of the reference
// (pillar_encoder.py:115-129 consumes exactly that), or -- fused with the voxelizer, SURVEY 8f.3 -- the voxelizer's
// per-voxel point-index lists (csrc/voxelize.cu `lists`: [batch][max_voxels][P] indices into `points`), in which case
// the [rows, P, ndim] tensor never exists.
struct PillarSrc {
  const float* voxels;        // dense mode
  const float* points;        // list mode: [n_total, ndim]
  const int* lists;           // list mode
  const int* counts;          // list mode: voxels per cloud [batch]
  int batch, max_voxels;
};

// stage the `cnt` points of output row `row` into `my` (one warp)
template <int NDIM_T>
__device__ __forceinline__ void stage_points(const PillarSrc& src, const int* pref, int row, int cnt, int P, int ndim_rt,
                                             int lane, float* my) {
  const int ndim = NDIM_T > 0 ? NDIM_T : ndim_rt;
  if (src.lists == nullptr) {
    const float* s = src.voxels + (size_t)row * P * ndim;
    for (int i = lane; i < cnt * ndim; i += 32) my[i] = s[i];
    return;
  }
  int b = 0;
  while (b + 1 < src.batch && row >= pref[b + 1]) ++b;
  const int* L = src.lists + ((size_t)b * src.max_voxels + (row - pref[b])) * P;
  for (int p = lane; p < cnt; p += 32) {
    const int idx = L[p];
    const float* q = src.points + (size_t)idx * ndim;
    if (NDIM_T == 4) {
      *reinterpret_cast<float4*>(my + p * 4) = idx < kListSentinel ? __ldg(reinterpret_cast<const float4*>(q)) : make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
      for (int d = 0; d < ndim; ++d) my[p * ndim + d] = idx < kListSentinel ? __ldg(q + d) : 0.f;
    }
  }
}

__global__ void __launch_bounds__(kPillarWarps * 32)
pillar_features_kernel(const PillarSrc src, const int* __restrict__ num_points,
                       const int* __restrict__ coors, const int* __restrict__ n_rows, int row_cap, int P,
                       int ndim, int units, const float* __restrict__ weight, const float* __restrict__ scale,
                       const float* __restrict__ shift, float vx, float vy, float x_offset, float y_offset,
                       float* __restrict__ out) {
  extern __shared__ float smem[];
  __shared__ int pref[kPillarMaxBatch + 1];
  const int n_in = ndim + 5;
  float* wt = smem;                                   // [n_in][units]  (transposed: lanes read consecutive floats)
  float* pts = smem + n_in * units;                   // [kPillarWarps][P * ndim]
  for (int i = threadIdx.x; i < n_in * units; i += blockDim.x) {
    const int c = i / n_in, f = i - c * n_in;         // weight is [units][n_in] (nn.Linear layout)
    wt[f * units + c] = weight[i];
  }
  if (threadIdx.x == 0 && src.lists != nullptr) {
    int acc = 0;
    for (int b = 0; b < src.batch; ++b) { pref[b] = acc; acc += src.counts[b]; }
    pref[src.batch] = acc;
  }
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* my = pts + (size_t)warp * P * ndim;
  const int n = min(*n_rows, row_cap);
  const int per_lane = units >> 5;

  float sc[kMaxUnitsPerLane], sh[kMaxUnitsPerLane];
#pragma unroll
  for (int j = 0; j < kMaxUnitsPerLane; ++j) {
    const bool on = j < per_lane;
    sc[j] = on ? scale[lane + 32 * j] : 0.f;
    sh[j] = on ? shift[lane + 32 * j] : 0.f;
  }

  for (int row = blockIdx.x * kPillarWarps + warp; row < row_cap; row += gridDim.x * kPillarWarps) {
    if (row >= n) {                                   // rows past the live count: defined (zero) output
      for (int j = 0; j < per_lane; ++j) out[(size_t)row * units + lane + 32 * j] = 0.f;
      continue;
    }
    const int cnt = min(max(num_points[row], 1), P);
    __syncwarp();
    stage_points<0>(src, pref, row, cnt, P, ndim, lane, my);
    __syncwarp();

    // per-pillar mean of xyz over the occupied slots (padding slots hold zeros in the reference's buffer, so
    // its sum over all P slots is the same sum)
    float sx = 0.f, sy = 0.f, sz = 0.f;
    for (int p = lane; p < cnt; p += 32) {
      sx += my[p * ndim + 0];
      sy += my[p * ndim + 1];
      sz += my[p * ndim + 2];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sx += __shfl_xor_sync(0xffffffffu, sx, o);
      sy += __shfl_xor_sync(0xffffffffu, sy, o);
      sz += __shfl_xor_sync(0xffffffffu, sz, o);
    }
    const float inv = 1.f / (float)num_points[row];
    const float mx = sx * inv, my_ = sy * inv, mz = sz * inv;
    const float cx = (float)coors[row * 4 + 3] * vx + x_offset;
    const float cy = (float)coors[row * 4 + 2] * vy + y_offset;

    // a padded slot contributes relu(shift) (all-zero features through Linear + BN), exactly as the reference's
    // masked tensor does; relu output is >= 0 so 0 is the neutral start when every slot is occupied
    float best[kMaxUnitsPerLane];
#pragma unroll
    for (int j = 0; j < kMaxUnitsPerLane; ++j) best[j] = (cnt < P) ? fmaxf(sh[j], 0.f) : 0.f;

    for (int p = 0; p < cnt; ++p) {
      float f[kMaxIn];
      const float* q = my + p * ndim;
      for (int d = 0; d < ndim; ++d) f[d] = q[d];
      f[ndim + 0] = q[0] - mx;
      f[ndim + 1] = q[1] - my_;
      f[ndim + 2] = q[2] - mz;
      f[ndim + 3] = q[0] - cx;
      f[ndim + 4] = q[1] - cy;
#pragma unroll
      for (int j = 0; j < kMaxUnitsPerLane; ++j) {
        if (j < per_lane) {
          float acc = 0.f;
          for (int d = 0; d < n_in; ++d) acc = fmaf(f[d], wt[d * units + lane + 32 * j], acc);
          best[j] = fmaxf(best[j], fmaxf(fmaf(acc, sc[j], sh[j]), 0.f));
        }
      }
    }
#pragma unroll
    for (int j = 0; j < kMaxUnitsPerLane; ++j)
      if (j < per_lane) out[(size_t)row * units + lane + 32 * j] = best[j];
  }
}

// ndim is a template parameter for the two point layouts the configs use so that f[] stays in registers
template <int NDIM>
__global__ void __launch_bounds__(kPillarWarps * 32)
pillar_features_fixed(const PillarSrc src, const int* __restrict__ num_points,
                      const int* __restrict__ coors, const int* __restrict__ n_rows, int row_cap, int P,
                      const float* __restrict__ weight, const float* __restrict__ scale,
                      const float* __restrict__ shift, float vx, float vy, float x_offset, float y_offset,
                      float* __restrict__ out) {
  constexpr int kIn = NDIM + 5, kUnits = 64;
  extern __shared__ float smem[];
  __shared__ int pref[kPillarMaxBatch + 1];
  float* wt = smem;                                   // [kIn][64]
  float* pts = smem + kIn * kUnits;
  for (int i = threadIdx.x; i < kIn * kUnits; i += blockDim.x) {
    const int c = i / kIn, f = i - c * kIn;
    wt[f * kUnits + c] = weight[i];
  }
  if (threadIdx.x == 0 && src.lists != nullptr) {
    int acc = 0;
    for (int b = 0; b < src.batch; ++b) { pref[b] = acc; acc += src.counts[b]; }
    pref[src.batch] = acc;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* my = pts + (size_t)warp * P * NDIM;
  const int n = min(*n_rows, row_cap);
  float w0[kIn], w1[kIn];
#pragma unroll
  for (int d = 0; d < kIn; ++d) { w0[d] = wt[d * kUnits + lane]; w1[d] = wt[d * kUnits + lane + 32]; }
  const float sc0 = scale[lane], sc1 = scale[lane + 32], sh0 = shift[lane], sh1 = shift[lane + 32];

  for (int row = blockIdx.x * kPillarWarps + warp; row < row_cap; row += gridDim.x * kPillarWarps) {
    if (row >= n) {
      out[(size_t)row * kUnits + lane] = 0.f;
      out[(size_t)row * kUnits + lane + 32] = 0.f;
      continue;
    }
    const int raw = num_points[row];
    const int cnt = min(max(raw, 1), P);
    __syncwarp();
    stage_points<NDIM>(src, pref, row, cnt, P, NDIM, lane, my);
    __syncwarp();
    float sx = 0.f, sy = 0.f, sz = 0.f;
    for (int p = lane; p < cnt; p += 32) {
      sx += my[p * NDIM + 0];
      sy += my[p * NDIM + 1];
      sz += my[p * NDIM + 2];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sx += __shfl_xor_sync(0xffffffffu, sx, o);
      sy += __shfl_xor_sync(0xffffffffu, sy, o);
      sz += __shfl_xor_sync(0xffffffffu, sz, o);
    }
    const float inv = 1.f / (float)raw;
    const float mx = sx * inv, my_ = sy * inv, mz = sz * inv;
    const float cx = (float)coors[row * 4 + 3] * vx + x_offset;
    const float cy = (float)coors[row * 4 + 2] * vy + y_offset;
    float b0 = (cnt < P) ? fmaxf(sh0, 0.f) : 0.f;
    float b1 = (cnt < P) ? fmaxf(sh1, 0.f) : 0.f;
    for (int p = 0; p < cnt; ++p) {
      float f[kIn];
      const float* q = my + p * NDIM;
#pragma unroll
      for (int d = 0; d < NDIM; ++d) f[d] = q[d];
      f[NDIM + 0] = f[0] - mx;
      f[NDIM + 1] = f[1] - my_;
      f[NDIM + 2] = f[2] - mz;
      f[NDIM + 3] = f[0] - cx;
      f[NDIM + 4] = f[1] - cy;
      float a0 = 0.f, a1 = 0.f;
#pragma unroll
      for (int d = 0; d < kIn; ++d) { a0 = fmaf(f[d], w0[d], a0); a1 = fmaf(f[d], w1[d], a1); }
      b0 = fmaxf(b0, fmaxf(fmaf(a0, sc0, sh0), 0.f));
      b1 = fmaxf(b1, fmaxf(fmaf(a1, sc1, sh1), 0.f));
    }
    out[(size_t)row * kUnits + lane] = b0;
    out[(size_t)row * kUnits + lane + 32] = b1;
  }
}

}  // namespace
}  // namespace d3b

using namespace d3b;

static int launch_pillars(const PillarSrc& src, const int32_t* num_points, const int32_t* coors, const int32_t* n_rows,
                          int32_t row_cap, int32_t max_points, int32_t ndim, int32_t units, const float* weight,
                          const float* scale, const float* shift, float vx, float vy, float x_offset, float y_offset,
                          float* out, void* stream_) {
  D3B_REQUIRE(num_points && coors && n_rows && weight && scale && shift && out, "pillar_features: null pointer");
  D3B_REQUIRE(row_cap >= 0 && max_points >= 1, "pillar_features: bad sizes (rows %d, points %d)", row_cap, max_points);
  if (ndim < 3 || ndim + 5 > kMaxIn || units < 32 || units > 32 * kMaxUnitsPerLane || (units & 31)) {
    set_error("pillar_features: ndim %d / units %d not built (3 <= ndim <= %d, units in {32,64,96,128})", ndim,
              units, kMaxIn - 5);
    return D3B_ERR_UNSUPPORTED;
  }
  if (row_cap == 0) return D3B_OK;
  cudaStream_t stream = (cudaStream_t)stream_;
  const size_t smem = ((size_t)(ndim + 5) * units + (size_t)kPillarWarps * max_points * ndim) * sizeof(float);
  D3B_REQUIRE(smem <= 200 * 1024, "pillar_features: %d points x %d dims do not fit in shared memory", max_points, ndim);
  const int grid = grid_for((long long)row_cap * 32, kPillarWarps * 32, 8);
#define D3B_PILLAR_FIXED(ND)                                                                                   \
  {                                                                                                            \
    if (smem > 48 * 1024)                                                                                      \
      D3B_CUDA(cudaFuncSetAttribute(pillar_features_fixed<ND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    pillar_features_fixed<ND><<<grid, kPillarWarps * 32, smem, stream>>>(                                      \
        src, num_points, coors, n_rows, row_cap, max_points, weight, scale, shift, vx, vy, x_offset,           \
        y_offset, out);                                                                                        \
  }
  if (units == 64 && ndim == 4) D3B_PILLAR_FIXED(4)
  else if (units == 64 && ndim == 5) D3B_PILLAR_FIXED(5)
  else {
    if (smem > 48 * 1024)
      D3B_CUDA(cudaFuncSetAttribute(pillar_features_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pillar_features_kernel<<<grid, kPillarWarps * 32, smem, stream>>>(
        src, num_points, coors, n_rows, row_cap, max_points, ndim, units, weight, scale, shift, vx, vy,
        x_offset, y_offset, out);
  }
#undef D3B_PILLAR_FIXED
  D3B_LAUNCH_CHECK();
  return D3B_OK;
}

extern "C" int d3b_pillar_features(const float* voxels, const int32_t* num_points, const int32_t* coors,
                                   const int32_t* n_rows, int32_t row_cap, int32_t max_points, int32_t ndim,
                                   int32_t units, const float* weight, const float* scale, const float* shift,
                                   float vx, float vy, float x_offset, float y_offset, float* out,
                                   void* stream_) {
  D3B_REQUIRE(voxels, "pillar_features: null voxels");
  PillarSrc src;
  src.voxels = voxels; src.points = nullptr; src.lists = nullptr; src.counts = nullptr; src.batch = 0; src.max_voxels = 0;
  return launch_pillars(src, num_points, coors, n_rows, row_cap, max_points, ndim, units, weight, scale, shift, vx, vy,
                        x_offset, y_offset, out, stream_);
}

extern "C" int d3b_pillar_features_lists(const float* points, const int32_t* lists, const int32_t* voxel_counts,
                                         int32_t batch, int32_t max_voxels, const int32_t* num_points,
                                         const int32_t* coors, const int32_t* n_rows, int32_t row_cap, int32_t max_points,
                                         int32_t ndim, int32_t units, const float* weight, const float* scale,
                                         const float* shift, float vx, float vy, float x_offset, float y_offset,
                                         float* out, void* stream_) {
  D3B_REQUIRE(points && lists && voxel_counts, "pillar_features_lists: null pointer");
  D3B_REQUIRE(batch >= 1 && batch <= kPillarMaxBatch && max_voxels >= 1, "pillar_features_lists: batch %d / max_voxels %d", batch, max_voxels);
  PillarSrc src;
  src.voxels = nullptr; src.points = points; src.lists = lists; src.counts = voxel_counts; src.batch = batch; src.max_voxels = max_voxels;
  return launch_pillars(src, num_points, coors, n_rows, row_cap, max_points, ndim, units, weight, scale, shift, vx, vy,
                        x_offset, y_offset, out, stream_);
}
