// Dense BEV convolutions (RPN blocks, deblocks, task heads) as implicit GEMMs on tcgen05, sm_100a.
//
// Reference: det3d/models/necks/rpn.py:82-159 (ZeroPad2d + Conv2d 3x3 [stride s] + BN + ReLU, n x (Conv2d 3x3 + BN +
// ReLU), deblock = Conv2d 1x1 / ConvTranspose2d(k = s, stride = s) + BN + ReLU, channel concat) and the 1x1 heads of
// det3d/models/bbox_heads/mg_head.py:198-230 -- fp32 cuDNN there.
//
// Round 1 ran these layers through the sparse gather kernel with a static "dense" rulebook: a dense tensor was
// re-gathered nine times through registers and st.shared, and the M128 x N128 SS-MMAs re-read both operands for each
// of the three split products -- the shared-memory pipe was the limiter (ncu: L1/shared 96 %, tensor 52 %).  Here:
//
//  * Activations are NHWC f16 planes (hi = f16(x), lo = f16(x - hi), see spconv16_sm100.cu), loaded by TMA with a 4-D
//    tensor map (C, W, H, B): box = 64 channels x 8 pixels x (16 + K - 1) rows, 128-byte swizzle, out-of-bounds
//    coordinates (the conv's zero padding, image borders, other samples) zero-filled by the TMA unit.  No thread
//    touches an activation on its way to the tensor core.
//  * Tile = 16 x 16 output pixels x C_out: two M = 128 halves (16 rows x 8 columns each) that share every weight
//    stage.  In a patch, eight consecutive 128-byte rows are one image row segment, so the matrix for kernel row ky is
//    the SAME staged patch with the UMMA descriptor start advanced by ky * 1024 bytes: one patch load per (kx, 64
//    channels) serves three (ky, kx) offsets.
//  * FP16x3: per k-step three MMAs (A_lo.B_hi, A_hi.B_lo, A_hi.B_hi) with fp32 accumulation in TMEM -- but only
//    across ONE A stage (the three kernel rows of a (kx, 64-channel slice)): the tensor core truncates on every
//    accumulation, and chaining all 216 MMAs of a 3x3x128 reduction into one accumulator shrank the outputs by ~4e-6
//    per layer (measured: error proportional to chain length), which 20 layers turn into > 1e-4.  A half's accumulator
//    therefore lives for one A stage (36 accumulations) in one of two TMEM buffers; eight accumulator warps add the
//    finished buffer into fp32 registers with round-to-nearest while the next A stage fills the other one, and at the
//    end of the tile apply bias / BN / ReLU and write the f16 planes.
//  * Warp roles: 1 TMA warp, 2 MMA warps (one issuing thread per half: a single thread issues one tcgen05.mma per
//    ~65 cycles -- clock-stamp trace -- which equals the pipe time of M128 x N128 x K16), 8 accumulator warps (one per
//    TMEM quadrant and half); A ring of 2 stages (4 patches each), B ring of 2..4 stages, all mbarrier-driven;
//    persistent grid (143 tiles at 200 x 176).
//  * stride 2 (RPN down-sampling blocks): the box is loaded with elementStrides = 2, one load per (ky, kx).
//  * groups: several weight blocks over the same input in one launch -- C_out = 256 as two N = 128 passes, and
//    ConvTranspose2d(k = s, stride = s) as s*s 1x1 convolutions whose epilogues write pixel (y*s + dy, x*s + dx)
//    into a channel slice of the concat buffer.
//
// Bound: tensor pipe (kind::f16).  fp32-equivalent flops per launch = 2 * B*H*W * K*K * C_in * C_out; the pipe
// executes 3x that.  Algorithmic bytes = B*H*W*(C_in + C_out)*4 + K*K*C_in*C_out*4.
#include <cuda.h>

#include "umma.cuh"

namespace d3b {

struct BvEpi {          // same fields as spconv16_sm100.cu (kept local: the two kernels are separate TUs)
  const float* bias;
  const float* scale;
  const float* shift;
  float acc_scale;
  int relu;
};

constexpr int kBvTileY = 16, kBvTileX = 16;     // output pixels per tile
constexpr int kBvHalfX = 8;                     // one M = 128 half = 16 rows x 8 columns
constexpr int kBvKc = 64;                       // channels per stage (128 bytes of f16)
// warp roles: TMA producer, one MMA-issuing warp per half (a single thread issues ~one tcgen05.mma per 65 cycles --
// measured -- which is exactly the pipe time of an M128 x N128 x K16 MMA: two issuers keep the pipe fed), 8 accumulator warps
constexpr int kBvTmaWarp = 0, kBvMmaWarp = 1, kBvEpiWarp0 = 3;
constexpr int kBvThreads = 32 * (kBvEpiWarp0 + 8);   // 352
constexpr int kBvAStages = 2;

template <int KS, int STRIDE, int COUT>
struct BvCfg {
  static constexpr int kPatchRows = STRIDE == 1 ? kBvTileY + KS - 1 : kBvTileY;      // rows in one staged patch
  static constexpr int kPatchBytes = kPatchRows * kBvHalfX * 128;
  static constexpr int kAStageBytes = 4 * kPatchBytes;                                // {half 0, half 1} x {hi, lo}
  static constexpr int kBBytes = 2 * COUT * 128;                                      // [B_hi rows | B_lo rows]
  static constexpr int kBStages = COUT >= 128 ? 2 : 4;
  static constexpr int kAccCols = COUT;                                               // one accumulator
  static constexpr int kTmemCols = 4 * kAccCols;                                      // {2 halves} x {2 buffers}: 128 .. 512
  static constexpr int kSmemBytes = kBvAStages * kAStageBytes + kBStages * kBBytes + 1024 + 256;
  // with stride 1 one A stage serves the KS kernel rows of a (kx, kb); with stride 2 every (ky, kx) has its own
  static constexpr int kRowsPerAStage = STRIDE == 1 ? KS : 1;
};

struct BvGeom {
  int batch, h_out, w_out;        // conv output grid
  int c_in, n_kb;
  int pad;
  int tiles_y, tiles_x;           // per sample
  int groups, cgroups;            // weight blocks; cgroups of them tile the output channels, the rest are (uy, ux)
  int up;                         // output pixel = (y*up + uy, x*up + ux)
  int out_h, out_w;               // output tensor grid (h_out*up, w_out*up)
  int out_channels, out_c0;       // row length of the output tensor and first channel written
  int seq;                        // launch counter of this translation unit (development traces only)
};

__device__ __forceinline__ uint32_t bv_pack_half2(__half a, __half b) {
  return (uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16);
}

template <int KS, int STRIDE, int COUT>
__global__ void __launch_bounds__(kBvThreads, 1)
bev_conv16_kernel(const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo, BvGeom g,
                  const __half* __restrict__ packed, BvEpi epi, __half* __restrict__ out_hi,
                  __half* __restrict__ out_lo, float* __restrict__ out_f32, int* __restrict__ overflow) {
  using Cfg = BvCfg<KS, STRIDE, COUT>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t a_base = smem_base;
  const uint32_t b_base = a_base + kBvAStages * Cfg::kAStageBytes;
  const uint32_t bar_base = b_base + Cfg::kBStages * Cfg::kBBytes;
  auto a_full = [&](int s) { return bar_base + 8u * s; };
  auto a_empty = [&](int s) { return bar_base + 8u * (kBvAStages + s); };
  auto b_full = [&](int s) { return bar_base + 8u * (2 * kBvAStages + s); };
  auto b_empty = [&](int s) { return bar_base + 8u * (2 * kBvAStages + Cfg::kBStages + s); };
  // accumulator hand-off, per (half, buffer): index half * 2 + buffer
  auto acc_full = [&](uint32_t hb) { return bar_base + 8u * (2 * kBvAStages + 2 * Cfg::kBStages + hb); };
  auto acc_empty = [&](uint32_t hb) { return bar_base + 8u * (2 * kBvAStages + 2 * Cfg::kBStages + 4 + hb); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_gen + (bar_base - smem_base) + 8 * (2 * kBvAStages + 2 * Cfg::kBStages + 8));

  D3B_CTA_MARK(0, g.seq);
  pdl_launch_dependents();           // the next kernel of the stream may start its prologue behind this one's tail
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_per_group = g.batch * g.tiles_y * g.tiles_x;
  const int n_tiles = tiles_per_group * g.groups;
  const int k_vol = KS * KS;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kBvAStages; ++s) { mbar_init(a_full(s), 1); mbar_init(a_empty(s), 2); }      // freed by both MMA threads
    for (int s = 0; s < Cfg::kBStages; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), 2); }
    for (uint32_t hb = 0; hb < 4; ++hb) {
      mbar_init(acc_full(hb), 1);      // tcgen05.commit of the half's MMAs of one A stage
      mbar_init(acc_empty(hb), 128);   // the half's four accumulator warps have read those partial sums
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    tma_prefetch_desc(&tm_hi);
    tma_prefetch_desc(&tm_lo);
  }
  if (warp == kBvMmaWarp) {     // (the first MMA warp owns the TMEM allocation)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)Cfg::kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;
  pdl_wait_prior_grid();             // everything below reads the previous layer's planes or writes buffers it may still read

  // tile -> (group, sample, y0, x0)
  auto decode = [&](int tile, int& grp, int& b, int& y0, int& x0) {
    grp = tile / tiles_per_group;
    int t = tile - grp * tiles_per_group;
    b = t / (g.tiles_y * g.tiles_x);
    t -= b * g.tiles_y * g.tiles_x;
    y0 = (t / g.tiles_x) * kBvTileY;
    x0 = (t % g.tiles_x) * kBvTileX;
  };

  if (warp == kBvTmaWarp) {
    // ===================== TMA producer (one elected lane) =====================
    if (lane == 0) {
      uint32_t a_it = 0, b_it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        int grp, b, y0, x0;
        decode(tile, grp, b, y0, x0);
        const __half* wgrp = packed + (size_t)grp * k_vol * g.n_kb * (Cfg::kBBytes / 2);
        for (int kb = 0; kb < g.n_kb; ++kb) {
          for (int kx = 0; kx < KS; ++kx) {
            for (int ky0 = 0; ky0 < KS; ky0 += Cfg::kRowsPerAStage) {
              // ---- A stage: {half 0, half 1} x {hi, lo} patches ----
              const int sa = a_it % kBvAStages;
              D3B_STAMP(0, a_it);
              D3B_WAIT(a_empty(sa), ((a_it / kBvAStages) & 1u) ^ 1u, 1);
              mbar_arrive_expect_tx(a_full(sa), Cfg::kAStageBytes);
              const uint32_t dst = a_base + sa * Cfg::kAStageBytes;
              const int cy = STRIDE == 1 ? y0 - g.pad : y0 * STRIDE + ky0 - g.pad;
#pragma unroll
              for (int half = 0; half < 2; ++half) {
                const int cx = (x0 + half * kBvHalfX) * STRIDE + kx - g.pad;
                tma_load_4d(dst + (2 * half) * Cfg::kPatchBytes, &tm_hi, kb * kBvKc, cx, cy, b, a_full(sa));
                tma_load_4d(dst + (2 * half + 1) * Cfg::kPatchBytes, &tm_lo, kb * kBvKc, cx, cy, b, a_full(sa));
              }
              D3B_STAMP(1, a_it);
              ++a_it;
              // ---- B stages: one per kernel row served by this A stage ----
              for (int r = 0; r < Cfg::kRowsPerAStage; ++r) {
                const int ky = ky0 + r;
                const int sb = b_it % Cfg::kBStages;
                D3B_STAMP(2, b_it);
                D3B_WAIT(b_empty(sb), ((b_it / Cfg::kBStages) & 1u) ^ 1u, 2);
                mbar_arrive_expect_tx(b_full(sb), Cfg::kBBytes);
                tma_bulk_g2s(b_base + sb * Cfg::kBBytes,
                             wgrp + ((size_t)(ky * KS + kx) * g.n_kb + kb) * (Cfg::kBBytes / 2), Cfg::kBBytes, b_full(sb));
                D3B_STAMP(3, b_it);
                ++b_it;
              }
            }
          }
        }
      }
    }
  } else if (warp < kBvEpiWarp0) {
    // ===================== MMA issuers: warp 1 -> half 0, warp 2 -> half 1 (one elected lane each) =====================
    // The half's accumulator lives for one A stage (the KS kernel rows of a (kx, 64-channel slice): 36 chained
    // accumulations -- the tensor core truncates on each, see the header) in one of two TMEM buffers; the accumulator warps
    // add it into fp32 registers while the next A stage fills the other buffer.
    constexpr uint32_t idesc = umma_idesc_f16(128, COUT);
    const int half = warp - kBvMmaWarp;
    uint32_t a_it = 0, b_it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      for (int kb = 0; kb < g.n_kb; ++kb) {
        const int n_ks = min(kBvKc / 16, (g.c_in - kb * kBvKc + 15) / 16);
        for (int kx = 0; kx < KS; ++kx) {
          for (int ky0 = 0; ky0 < KS; ky0 += Cfg::kRowsPerAStage) {
            const int sa = a_it % kBvAStages;
            const uint32_t hb = half * 2 + (a_it & 1u);
            if (lane == 0 && half == 0) D3B_STAMP(4, a_it);
            D3B_WAIT(acc_empty(hb), ((a_it >> 1) & 1u) ^ 1u, 3);        // this buffer's previous sums have been read out
            D3B_WAIT(a_full(sa), (a_it / kBvAStages) & 1u, 4);
            if (lane == 0 && half == 0) D3B_STAMP(5, a_it);
            const uint32_t d_addr = tmem_d + hb * Cfg::kAccCols;
            for (int r = 0; r < Cfg::kRowsPerAStage; ++r, ++b_it) {
              const int sb = b_it % Cfg::kBStages;
              D3B_WAIT(b_full(sb), (b_it / Cfg::kBStages) & 1u, 5);
              if (lane == 0 && half == 0) D3B_STAMP(6, b_it);
              tc_fence_after();
              {
                // warp-uniform instruction stream (descriptor arithmetic stays in uniform registers); lane 0 issues
                const uint32_t issue = lane == 0 ? 1u : 0u;
                const uint32_t b_hi = b_base + sb * Cfg::kBBytes;
                // stride 1: kernel row r of the staged patch = the same bytes 8 rows (1024 B) further down
                const uint32_t row_adv = STRIDE == 1 ? (uint32_t)(ky0 + r) * 1024u : 0u;
                const uint32_t ah = a_base + sa * Cfg::kAStageBytes + (2 * half) * Cfg::kPatchBytes + row_adv;
                const uint64_t da_hi = umma_desc_sw128(ah), da_lo = umma_desc_sw128(ah + Cfg::kPatchBytes);
                const uint64_t db_hi = umma_desc_sw128(b_hi), db_lo = umma_desc_sw128(b_hi + COUT * 128);
#pragma unroll
                for (int ks = 0; ks < kBvKc / 16; ++ks) {
                  if (ks < n_ks) {
                    const uint64_t adv = (uint64_t)(ks * 2);     // 32 bytes along K = 2 descriptor address units
                    // small terms first, the dominant hi.hi product last
                    tc_mma_f16_if(issue, d_addr, da_lo + adv, db_hi + adv, idesc, (r | ks) ? 1u : 0u);
                    tc_mma_f16_if(issue, d_addr, da_hi + adv, db_lo + adv, idesc, 1u);
                    tc_mma_f16_if(issue, d_addr, da_hi + adv, db_hi + adv, idesc, 1u);
                  }
                }
                tc_commit_if(issue, b_empty(sb));
                if (lane == 0 && half == 0) D3B_STAMP(7, b_it);
              }
            }
            tc_commit_if(lane == 0 ? 1u : 0u, a_empty(sa));
            tc_commit_if(lane == 0 ? 1u : 0u, acc_full(hb));
            ++a_it;
          }
        }
      }
    }
  } else {
    // ===================== accumulator warps: TMEM partial sums -> fp32 registers -> bias / BN / ReLU -> NHWC planes =====
    const int ew = warp - kBvEpiWarp0;          // 0..7
    const int quad = warp & 3;                  // TMEM lane quadrant this warp may read (warps 3..6 and 7..10: 3,0,1,2)
    const int half = ew >> 2;
    const int m = quad * 32 + lane;             // pixel of the half: row m / 8, column m % 8
    const int n_groups = g.n_kb * k_vol / Cfg::kRowsPerAStage;      // accumulator groups (A stages) per tile
    bool ovf = false;
    uint32_t a_it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      int grp, b, y0, x0;
      decode(tile, grp, b, y0, x0);
      float acc[COUT];
#pragma unroll
      for (int q = 0; q < COUT; ++q) acc[q] = 0.f;
#pragma unroll 1
      for (int gi = 0; gi < n_groups; ++gi, ++a_it) {
        const uint32_t hb = half * 2 + (a_it & 1u);
        if (ew == 0 && lane == 0) D3B_STAMP(8, a_it);
        D3B_WAIT(acc_full(hb), (a_it >> 1) & 1u, 6);
        if (ew == 0 && lane == 0) D3B_STAMP(9, a_it);
        tc_fence_after();
        const uint32_t t0 = tmem_d + hb * Cfg::kAccCols + ((uint32_t)(quad * 32) << 16);
        const int kb_g = gi * Cfg::kRowsPerAStage / k_vol;
        const int n_chain = 3 * Cfg::kRowsPerAStage * min(kBvKc / 16, (g.c_in - kb_g * kBvKc + 15) / 16);   // MMAs chained into the buffer
        const float f_fix = 1.f + (float)n_chain * kTruncLossPerMma;
#pragma unroll
        for (int c0 = 0; c0 < COUT; c0 += 32) {
          uint32_t r0[16], r1[16];
          tc_ld16_nowait(t0 + c0, r0);
          tc_ld16_nowait(t0 + c0 + 16, r1);
          tc_ld_wait();
#pragma unroll
          for (int q = 0; q < 16; ++q) acc[c0 + q] = fmaf(__uint_as_float(r0[q]), f_fix, acc[c0 + q]);
#pragma unroll
          for (int q = 0; q < 16; ++q) acc[c0 + 16 + q] = fmaf(__uint_as_float(r1[q]), f_fix, acc[c0 + 16 + q]);
        }
        tc_fence_before();
        mbar_arrive(acc_empty(hb));
        if (ew == 0 && lane == 0) D3B_STAMP(10, a_it);
      }
      const int y = y0 + (m >> 3), x = x0 + half * kBvHalfX + (m & 7);
      if (!(y < g.h_out && x < g.w_out)) continue;
      const int cg = grp % g.cgroups, ug = grp / g.cgroups;
      const int oy = y * g.up + ug / g.up, ox = x * g.up + ug % g.up;
      const size_t row_off = (((size_t)b * g.out_h + oy) * g.out_w + ox) * (size_t)g.out_channels + g.out_c0 + cg * COUT;
      const int pcol = grp * COUT;              // per-group epilogue parameters are laid out group-major
#pragma unroll
      for (int c0 = 0; c0 < COUT; c0 += 16) {
        float* v = acc + c0;                      // in place: the running sums of this chunk are dead afterwards
#pragma unroll
        for (int q = 0; q < 16; ++q) v[q] *= epi.acc_scale;
        if (epi.bias) {
#pragma unroll
          for (int q = 0; q < 16; q += 4) {
            const float4 bb = __ldg(reinterpret_cast<const float4*>(epi.bias + pcol + c0 + q));
            v[q] += bb.x; v[q + 1] += bb.y; v[q + 2] += bb.z; v[q + 3] += bb.w;
          }
        }
        if (epi.scale) {
#pragma unroll
          for (int q = 0; q < 16; q += 4) {
            const float4 sc = __ldg(reinterpret_cast<const float4*>(epi.scale + pcol + c0 + q));
            const float4 sh = __ldg(reinterpret_cast<const float4*>(epi.shift + pcol + c0 + q));
            v[q] = fmaf(v[q], sc.x, sh.x); v[q + 1] = fmaf(v[q + 1], sc.y, sh.y);
            v[q + 2] = fmaf(v[q + 2], sc.z, sh.z); v[q + 3] = fmaf(v[q + 3], sc.w, sh.w);
          }
        }
        if (epi.relu) {
#pragma unroll
          for (int q = 0; q < 16; ++q) v[q] = fmaxf(v[q], 0.f);
        }
        if (out_hi) {
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int q = 0; q < 16; q += 2) {
            __half h0, l0, h1, l1;
            split_f16(v[q], h0, l0);
            split_f16(v[q + 1], h1, l1);
            hi[q >> 1] = bv_pack_half2(h0, h1);
            lo[q >> 1] = bv_pack_half2(l0, l1);
            ovf |= !(fabsf(v[q]) < 65504.f) | !(fabsf(v[q + 1]) < 65504.f);
          }
          uint4* ph = reinterpret_cast<uint4*>(out_hi + row_off + c0);
          uint4* pl = reinterpret_cast<uint4*>(out_lo + row_off + c0);
          ph[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          ph[1] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
          pl[0] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          pl[1] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
        }
        if (out_f32) {
          float4* pf = reinterpret_cast<float4*>(out_f32 + row_off + c0);
#pragma unroll
          for (int q = 0; q < 16; q += 4) pf[q >> 2] = make_float4(v[q], v[q + 1], v[q + 2], v[q + 3]);
        }
      }
    }
    if (ovf && overflow) atomicOr(overflow, 1);
  }

  tc_fence_before();
  __syncthreads();
  D3B_CTA_MARK(1, g.seq);
  if (warp == kBvMmaWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"((uint32_t)Cfg::kTmemCols)
                 : "memory");
  }
}

// ======================================================================================================================
// Variant "channel-stationary" (3x3, stride 1, C_out = 128 per group, C_in % 64 == 0, up = 1): the GEMM is transposed,
//     D^T[C_out = 128 (M, TMEM lanes), 256 pixels (N, TMEM columns)] += W[C_out, K16] . Act[256 pixels, K16]^T
// so ONE N = 256 MMA covers the whole 16 x 16 pixel tile.  The kernel above runs the tile as two M = 128 halves: per
// k-step six M128 x N128 MMAs re-read 48 KB of shared memory (A 4 KB + B 4 KB each) for 384 cycles of math, and the
// shared-memory pipe (which also takes the TMA writes) caps the tensor pipe at ~45 % (ncu, clock-stamp traces).  Here
// a k-step is three M128 x N256 MMAs: 36 KB of operand reads for the same 384 cycles, and half the instructions to issue
// (one issuing thread suffices).  Same products, same order, same chain length (36) and truncation correction.
//  * Activations: one TMA box per plane = 64 channels x 16 pixels x 18 rows: an image row is 2048 B (two swizzle groups),
//    pixel (r, x) of the tile is operand row r * 16 + x, kernel row ky = descriptor start + ky * 2048.
//  * Weights: the packed [W_hi | W_lo] image is K-major SW128 with C_out rows -- byte for byte usable as the A operand.
//  * TMEM: 2 buffers x 256 columns; lane = output channel.  Eight accumulator warps: warp % 4 = lane quadrant (32
//    channels), (warp - 2) / 4 = column half (128 pixels = 8 tile rows); a thread keeps 128 running sums of ONE channel,
//    so bias / BN parameters are per-thread scalars.
//  * Stores: the planes are NHWC, a thread holds one channel of 128 pixels.  Groups of 4 lanes transpose 4 pixels x 4
//    channels of packed f16 with two warp shuffles per plane, then every lane writes 8 bytes (4 channels of one pixel):
//    a warp store covers 4 pixels x 64 contiguous bytes.
constexpr int kBwThreads = 352;                                   // warp 0 activation TMA, 1 MMA, 2..9 accumulators, 10 weights
constexpr int kBwEpiWarp0 = 2;
constexpr int kBwWeightWarp = 10;
constexpr int kBwCout = 128;
constexpr int kBwPatchRows = kBvTileY + 2;                        // 18
constexpr int kBwRowBytes = kBvTileX * 128;                       // one image row of the patch: 16 pixels x 128 B
constexpr int kBwPatchBytes = kBwPatchRows * kBwRowBytes;         // 36864
constexpr int kBwAStageBytes = 2 * kBwPatchBytes;                 // hi, lo
constexpr int kBwBBytes = 2 * kBwCout * 128;                      // [W_hi rows | W_lo rows]
constexpr int kBwBStages = 2;
constexpr int kBwAccCols = kBvTileY * kBvTileX;                   // 256 pixel columns per buffer
constexpr int kBwTmemCols = 2 * kBwAccCols;                       // 512
constexpr int kBwSmemBytes = kBvAStages * kBwAStageBytes + kBwBStages * kBwBBytes + 1024 + 256;
static_assert(kBvAStages == 2, "the accumulator buffer index is the A stage index");

__global__ void __launch_bounds__(kBwThreads, 1)
bev_conv16_cs_kernel(const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo, BvGeom g,
                     const __half* __restrict__ packed, BvEpi epi, __half* __restrict__ out_hi,
                     __half* __restrict__ out_lo, float* __restrict__ out_f32, int* __restrict__ overflow) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t a_base = smem_base;
  const uint32_t b_base = a_base + kBvAStages * kBwAStageBytes;
  const uint32_t bar_base = b_base + kBwBStages * kBwBBytes;
  auto a_full = [&](int s) { return bar_base + 8u * s; };
  auto a_empty = [&](int s) { return bar_base + 8u * (2 + s); };
  auto b_full = [&](int s) { return bar_base + 8u * (4 + s); };
  auto b_empty = [&](int s) { return bar_base + 8u * (6 + s); };
  auto acc_full = [&](uint32_t buf) { return bar_base + 8u * (8 + buf); };
  auto acc_empty = [&](uint32_t buf) { return bar_base + 8u * (10 + buf); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_gen + (bar_base - smem_base) + 8 * 12);

  D3B_CTA_MARK(0, g.seq);
  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_per_group = g.batch * g.tiles_y * g.tiles_x;
  const int n_tiles = tiles_per_group * g.groups;

  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(a_full(s), 1); mbar_init(a_empty(s), 1);
      mbar_init(b_full(s), 1); mbar_init(b_empty(s), 1);
      mbar_init(acc_full(s), 1);       // tcgen05.commit after the 36 MMAs of one A stage
      mbar_init(acc_empty(s), 256);    // all eight accumulator warps have read the buffer
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    tma_prefetch_desc(&tm_hi);
    tma_prefetch_desc(&tm_lo);
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)kBwTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;
  // (griddepcontrol.wait is executed by the threads that touch the previous layer's planes or write outputs: the activation
  // producer and the accumulator warps; weight staging and the MMA issue run ahead of the previous grid's tail)

  auto decode = [&](int tile, int& grp, int& b, int& y0, int& x0) {
    grp = tile / tiles_per_group;
    int t = tile - grp * tiles_per_group;
    b = t / (g.tiles_y * g.tiles_x);
    t -= b * g.tiles_y * g.tiles_x;
    y0 = (t / g.tiles_x) * kBvTileY;
    x0 = (t % g.tiles_x) * kBvTileX;
  };

  if (warp == 0) {
    // ===================== activation producer (TMA tensor loads) =====================
    // (Weights have their own producer warp: with one thread feeding both rings, the next patch load queued up behind a
    // wait for a free weight stage and every A stage began with a ~2000-cycle bubble -- clock-stamp trace.)
    if (lane == 0) {
      pdl_wait_prior_grid();           // the planes are the previous layer's output
      uint32_t a_it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        int grp, b, y0, x0;
        decode(tile, grp, b, y0, x0);
        for (int kb = 0; kb < g.n_kb; ++kb) {
          for (int kx = 0; kx < 3; ++kx, ++a_it) {
            const int sa = a_it & 1;
            D3B_STAMP(0, a_it);
            D3B_WAIT(a_empty(sa), ((a_it >> 1) & 1u) ^ 1u, 1);
            mbar_arrive_expect_tx(a_full(sa), kBwAStageBytes);
            const uint32_t dst = a_base + sa * kBwAStageBytes;
            tma_load_4d(dst, &tm_hi, kb * kBvKc, x0 + kx - g.pad, y0 - g.pad, b, a_full(sa));
            tma_load_4d(dst + kBwPatchBytes, &tm_lo, kb * kBvKc, x0 + kx - g.pad, y0 - g.pad, b, a_full(sa));
            D3B_STAMP(1, a_it);
          }
        }
      }
    }
  } else if (warp == kBwWeightWarp) {
    // ===================== weight producer (bulk copies; weights do not depend on the previous grid) =====================
    if (lane == 0) {
      uint32_t b_it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int grp = tile / tiles_per_group;
        const __half* wgrp = packed + (size_t)grp * 9 * g.n_kb * (kBwBBytes / 2);
        for (int kb = 0; kb < g.n_kb; ++kb) {
          for (int kx = 0; kx < 3; ++kx) {
            for (int ky = 0; ky < 3; ++ky, ++b_it) {
              const int sb = b_it & 1;
              D3B_STAMP(2, b_it);
              D3B_WAIT(b_empty(sb), ((b_it >> 1) & 1u) ^ 1u, 2);
              mbar_arrive_expect_tx(b_full(sb), kBwBBytes);
              tma_bulk_g2s(b_base + sb * kBwBBytes, wgrp + ((size_t)(ky * 3 + kx) * g.n_kb + kb) * (kBwBBytes / 2),
                           kBwBBytes, b_full(sb));
              D3B_STAMP(3, b_it);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (warp-uniform stream, lane 0 issues) =====================
    constexpr uint32_t idesc = umma_idesc_f16(kBwCout, kBwAccCols);     // M = 128 channels, N = 256 pixels
    const uint32_t issue = lane == 0 ? 1u : 0u;
    uint32_t a_it = 0, b_it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      for (int kb = 0; kb < g.n_kb; ++kb) {
        for (int kx = 0; kx < 3; ++kx, ++a_it) {
          const uint32_t sa = a_it & 1u;                                 // A stage == accumulator buffer
          if (lane == 0) D3B_STAMP(4, a_it);
          D3B_WAIT(acc_empty(sa), ((a_it >> 1) & 1u) ^ 1u, 3);
          D3B_WAIT(a_full(sa), (a_it >> 1) & 1u, 4);
          if (lane == 0) D3B_STAMP(5, a_it);
          const uint32_t d_addr = tmem_d + sa * kBwAccCols;
          for (int ky = 0; ky < 3; ++ky, ++b_it) {
            const uint32_t sb = b_it & 1u;
            D3B_WAIT(b_full(sb), (b_it >> 1) & 1u, 5);
            if (lane == 0) D3B_STAMP(6, b_it);
            tc_fence_after();
            const uint32_t w_hi = b_base + sb * kBwBBytes;
            const uint32_t x_hi = a_base + sa * kBwAStageBytes + (uint32_t)ky * kBwRowBytes;
            const uint64_t dw_hi = umma_desc_sw128(w_hi), dw_lo = umma_desc_sw128(w_hi + kBwCout * 128);
            const uint64_t dx_hi = umma_desc_sw128(x_hi), dx_lo = umma_desc_sw128(x_hi + kBwPatchBytes);
#pragma unroll
            for (int ks = 0; ks < kBvKc / 16; ++ks) {
              const uint64_t adv = (uint64_t)(ks * 2);
              // small terms first, the dominant hi.hi product last (the order of the pixel-stationary kernel)
              tc_mma_f16_if(issue, d_addr, dw_hi + adv, dx_lo + adv, idesc, (ky | ks) ? 1u : 0u);
              tc_mma_f16_if(issue, d_addr, dw_lo + adv, dx_hi + adv, idesc, 1u);
              tc_mma_f16_if(issue, d_addr, dw_hi + adv, dx_hi + adv, idesc, 1u);
            }
            tc_commit_if(issue, b_empty(sb));
            if (lane == 0) D3B_STAMP(7, b_it);
          }
          tc_commit_if(issue, a_empty(sa));
          tc_commit_if(issue, acc_full(sa));
        }
      }
    }
  } else {
    // ===================== accumulator warps =====================
    pdl_wait_prior_grid();                              // (output buffers may still be read by earlier kernels)
    const int quad = warp & 3;                          // TMEM lane quadrant (warps 2..9: 2,3,0,1,2,3,0,1)
    const int colhalf = (warp - kBwEpiWarp0) >> 2;      // pixels [128 * colhalf, +128) = tile rows [8 * colhalf, +8)
    const int ch = quad * 32 + lane;                    // output channel of the group
    const int n_groups = g.n_kb * 3;
    const float f_fix = 1.f + 36.f * kTruncLossPerMma;  // 3 kernel rows x 4 k-steps x 3 products chained per buffer
    bool ovf = false;
    uint32_t a_it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      int grp, b, y0, x0;
      decode(tile, grp, b, y0, x0);
      float acc[128];
#pragma unroll
      for (int q = 0; q < 128; ++q) acc[q] = 0.f;
#pragma unroll 1
      for (int gi = 0; gi < n_groups; ++gi, ++a_it) {
        const uint32_t buf = a_it & 1u;
        if (warp == kBwEpiWarp0 && lane == 0) D3B_STAMP(8, a_it);
        D3B_WAIT(acc_full(buf), (a_it >> 1) & 1u, 6);
        if (warp == kBwEpiWarp0 && lane == 0) D3B_STAMP(9, a_it);
        tc_fence_after();
        const uint32_t t0 = tmem_d + buf * kBwAccCols + colhalf * 128 + ((uint32_t)(quad * 32) << 16);
#pragma unroll
        for (int c0 = 0; c0 < 128; c0 += 32) {
          uint32_t r0[16], r1[16];
          tc_ld16_nowait(t0 + c0, r0);
          tc_ld16_nowait(t0 + c0 + 16, r1);
          tc_ld_wait();
#pragma unroll
          for (int q = 0; q < 16; ++q) acc[c0 + q] = fmaf(__uint_as_float(r0[q]), f_fix, acc[c0 + q]);
#pragma unroll
          for (int q = 0; q < 16; ++q) acc[c0 + 16 + q] = fmaf(__uint_as_float(r1[q]), f_fix, acc[c0 + 16 + q]);
        }
        tc_fence_before();
        mbar_arrive(acc_empty(buf));
        if (warp == kBwEpiWarp0 && lane == 0) D3B_STAMP(10, a_it);
      }
      // ---- epilogue: per-channel scalars, 4 x 4 (pixel x channel) transposes through shuffles, 8-byte stores ----
      const int cg = grp % g.cgroups;
      const int pcol = grp * kBwCout + ch;
      const float e_bias = epi.bias ? __ldg(epi.bias + pcol) : 0.f;
      const float e_scale = epi.scale ? __ldg(epi.scale + pcol) : 1.f;
      const float e_shift = epi.scale ? __ldg(epi.shift + pcol) : 0.f;
      const size_t chan0 = (size_t)g.out_c0 + (size_t)cg * kBwCout;
      const bool even = (lane & 1) == 0, low2 = (lane & 2) == 0;
#pragma unroll
      for (int q0 = 0; q0 < 128; q0 += 4) {
        const int y = y0 + colhalf * 8 + (q0 >> 4);
        const int xb = x0 + (q0 & 15);
        uint32_t eh[4], el[4];                            // 16-bit patterns of the four pixels (this lane's channel)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float v = acc[q0 + j] * epi.acc_scale;
          if (epi.bias) v += e_bias;
          if (epi.scale) v = fmaf(v, e_scale, e_shift);
          if (epi.relu) v = fmaxf(v, 0.f);
          acc[q0 + j] = v;
          __half h, l;
          split_f16(v, h, l);
          eh[j] = (uint32_t)__half_as_ushort(h);
          el[j] = (uint32_t)__half_as_ushort(l);
          ovf |= !(fabsf(v) < 65504.f);
        }
        if (out_f32 && y < g.h_out) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (xb + j < g.w_out)
              out_f32[(((size_t)b * g.out_h + y) * g.out_w + xb + j) * (size_t)g.out_channels + chan0 + ch] = acc[q0 + j];
        }
        if (out_hi) {
          // stage A (partner lane ^ 1): even lanes collect pixels 0 and 2, odd lanes pixels 1 and 3, for a channel pair
          const uint32_t sh_h = even ? (eh[1] | (eh[3] << 16)) : (eh[0] | (eh[2] << 16));
          const uint32_t sh_l = even ? (el[1] | (el[3] << 16)) : (el[0] | (el[2] << 16));
          const uint32_t rh = __shfl_xor_sync(0xffffffffu, sh_h, 1), rl = __shfl_xor_sync(0xffffffffu, sh_l, 1);
          uint32_t ua_h, ub_h, ua_l, ub_l;                // ua: pixel (lane & 1), ub: pixel 2 + (lane & 1); low half = lower channel
          if (even) {
            ua_h = eh[0] | (rh << 16);           ub_h = eh[2] | (rh & 0xffff0000u);
            ua_l = el[0] | (rl << 16);           ub_l = el[2] | (rl & 0xffff0000u);
          } else {
            ua_h = (rh & 0xffffu) | (eh[1] << 16); ub_h = (rh >> 16) | (eh[3] << 16);
            ua_l = (rl & 0xffffu) | (el[1] << 16); ub_l = (rl >> 16) | (el[3] << 16);
          }
          // stage B (partner lane ^ 2): lanes 0,1 of a quad keep pixels 0,1 and take the upper channel pair; lanes 2,3 keep 2,3
          const uint32_t th = __shfl_xor_sync(0xffffffffu, low2 ? ub_h : ua_h, 2);
          const uint32_t tl = __shfl_xor_sync(0xffffffffu, low2 ? ub_l : ua_l, 2);
          const uint2 wh = low2 ? make_uint2(ua_h, th) : make_uint2(th, ub_h);
          const uint2 wl = low2 ? make_uint2(ua_l, tl) : make_uint2(tl, ub_l);
          const int x = xb + (lane & 3);
          if (y < g.h_out && x < g.w_out) {
            const size_t off = (((size_t)b * g.out_h + y) * g.out_w + x) * (size_t)g.out_channels + chan0 + quad * 32 + (lane & ~3);
            *reinterpret_cast<uint2*>(out_hi + off) = wh;
            *reinterpret_cast<uint2*>(out_lo + off) = wl;
          }
        }
      }
    }
    if (warp == kBwEpiWarp0 && lane == 0) D3B_STAMP(11, 0);
    if (ovf && overflow) atomicOr(overflow, 1);
  }

  tc_fence_before();
  __syncthreads();
  D3B_CTA_MARK(1, g.seq);
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"((uint32_t)kBwTmemCols)
                 : "memory");
  }
}

// ---- host side: tensor maps through the driver entry point (libcuda is not linked: CPU hosts must dlopen us) ----------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
  static std::atomic<void*> cached{nullptr};
  void* fn = cached.load(std::memory_order_acquire);
  if (fn == nullptr) {
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    cached.store(fn, std::memory_order_release);
  }
  return (EncodeTiledFn)fn;
}

// planes [B, H, W, C] f16 -> map with box (64 channels, box_w pixels, box_h rows, 1 sample), 128B swizzle, zero OOB fill
static int make_map(CUtensorMap* map, const void* base, int batch, int h, int w, int c, int box_w, int box_h, int stride) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (enc == nullptr) {
    set_error("cuTensorMapEncodeTiled is not available from this driver");
    return D3B_ERR_CUDA;
  }
  const cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)batch};
  const cuuint64_t strides[3] = {(cuuint64_t)c * 2, (cuuint64_t)w * c * 2, (cuuint64_t)h * w * c * 2};
  // with a traversal stride s the box spans (n - 1) * s + 1 tensor elements to deliver n of them
  const cuuint32_t box[4] = {(cuuint32_t)kBvKc, (cuuint32_t)((box_w - 1) * stride + 1), (cuuint32_t)((box_h - 1) * stride + 1), 1u};
  const cuuint32_t estr[4] = {1u, (cuuint32_t)stride, (cuuint32_t)stride, 1u};
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) for planes [%d,%d,%d,%d] box [%d,%d] stride %d", (int)r, batch, h, w, c,
              box_w, box_h, stride);
    return D3B_ERR_CUDA;
  }
  return D3B_OK;
}

template <int KS, int STRIDE, int COUT>
static int launch_bev(const d3b_bev16_params* p, const BvGeom& g, cudaStream_t stream) {
  using Cfg = BvCfg<KS, STRIDE, COUT>;
  static SmemOptIn optin;
  D3B_CUDA(ensure_dynamic_smem(bev_conv16_kernel<KS, STRIDE, COUT>, Cfg::kSmemBytes, optin));
  CUtensorMap tm_hi, tm_lo;
  int st = make_map(&tm_hi, p->in_hi, p->batch, p->h_in, p->w_in, p->c_in, kBvHalfX, Cfg::kPatchRows, STRIDE);
  if (st != D3B_OK) return st;
  st = make_map(&tm_lo, p->in_lo, p->batch, p->h_in, p->w_in, p->c_in, kBvHalfX, Cfg::kPatchRows, STRIDE);
  if (st != D3B_OK) return st;
  BvEpi e;
  e.bias = p->bias; e.scale = p->scale; e.shift = p->shift; e.acc_scale = p->acc_scale; e.relu = p->relu;
  const int n_tiles = g.batch * g.tiles_y * g.tiles_x * g.groups;
  const int grid = n_tiles < kNumSMs ? n_tiles : kNumSMs;
  D3B_CUDA(launch_maybe_pdl(bev_conv16_kernel<KS, STRIDE, COUT>, dim3(grid), dim3(kBvThreads), Cfg::kSmemBytes, stream, tm_hi,
                            tm_lo, g, (const __half*)p->weight_packed, e, (__half*)p->out_hi, (__half*)p->out_lo, p->out_f32,
                            (int*)p->overflow));
  D3B_LAUNCH_CHECK();
  return D3B_OK;
}

// channel-stationary variant: 3x3, stride 1, one or more output blocks of exactly 128 channels, no sub-pixel groups
static int launch_bev_cs(const d3b_bev16_params* p, const BvGeom& g, cudaStream_t stream) {
  static SmemOptIn optin;
  D3B_CUDA(ensure_dynamic_smem(bev_conv16_cs_kernel, kBwSmemBytes, optin));
  CUtensorMap tm_hi, tm_lo;
  int st = make_map(&tm_hi, p->in_hi, p->batch, p->h_in, p->w_in, p->c_in, kBvTileX, kBwPatchRows, 1);
  if (st != D3B_OK) return st;
  st = make_map(&tm_lo, p->in_lo, p->batch, p->h_in, p->w_in, p->c_in, kBvTileX, kBwPatchRows, 1);
  if (st != D3B_OK) return st;
  BvEpi e;
  e.bias = p->bias; e.scale = p->scale; e.shift = p->shift; e.acc_scale = p->acc_scale; e.relu = p->relu;
  const int n_tiles = g.batch * g.tiles_y * g.tiles_x * g.groups;
  const int grid = n_tiles < kNumSMs ? n_tiles : kNumSMs;
  D3B_CUDA(launch_maybe_pdl(bev_conv16_cs_kernel, dim3(grid), dim3(kBwThreads), kBwSmemBytes, stream, tm_hi, tm_lo, g,
                            (const __half*)p->weight_packed, e, (__half*)p->out_hi, (__half*)p->out_lo, p->out_f32,
                            (int*)p->overflow));
  D3B_LAUNCH_CHECK();
  return D3B_OK;
}

}  // namespace d3b

using namespace d3b;

extern "C" int d3b_bev_conv16(const d3b_bev16_params* p, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  D3B_REQUIRE(p && p->in_hi && p->in_lo && p->weight_packed, "d3b_bev_conv16: null argument");
  D3B_REQUIRE(p->batch >= 1 && p->h_in >= 1 && p->w_in >= 1 && p->c_in >= 16 && p->c_in % 16 == 0,
              "d3b_bev_conv16: bad input shape [%d,%d,%d,%d] (C_in must be a multiple of 16)", p->batch, p->h_in, p->w_in, p->c_in);
  D3B_REQUIRE((p->ksize == 1 || p->ksize == 3) && (p->stride == 1 || p->stride == 2) && !(p->ksize == 1 && p->stride != 1),
              "d3b_bev_conv16: ksize %d stride %d not built (3x3 s1/s2, 1x1 s1)", p->ksize, p->stride);
  D3B_REQUIRE(p->pad >= 0 && p->pad <= p->ksize / 2 + 1, "d3b_bev_conv16: pad %d", p->pad);
  D3B_REQUIRE(p->up >= 1 && p->up <= 4 && p->cgroups >= 1 && p->groups == p->cgroups * p->up * p->up,
              "d3b_bev_conv16: groups %d != cgroups %d * up^2 (up %d)", p->groups, p->cgroups, p->up);
  D3B_REQUIRE((p->out_hi != nullptr) == (p->out_lo != nullptr) && (p->out_hi || p->out_f32),
              "d3b_bev_conv16: give out_hi + out_lo and/or out_f32");
  D3B_REQUIRE((p->scale == nullptr) == (p->shift == nullptr), "d3b_bev_conv16: scale and shift go together");
  D3B_REQUIRE(p->out_channels % 8 == 0 && p->out_c0 % 8 == 0 && p->out_c0 + p->cgroups * p->c_out <= p->out_channels,
              "d3b_bev_conv16: output channel slice [%d, %d) does not fit rows of %d", p->out_c0,
              p->out_c0 + p->cgroups * p->c_out, p->out_channels);
  BvGeom g;
  g.batch = p->batch;
  g.h_out = (p->h_in + 2 * p->pad - p->ksize) / p->stride + 1;
  g.w_out = (p->w_in + 2 * p->pad - p->ksize) / p->stride + 1;
  D3B_REQUIRE(g.h_out >= 1 && g.w_out >= 1, "d3b_bev_conv16: empty output grid");
  g.c_in = p->c_in;
  g.n_kb = (p->c_in + kBvKc - 1) / kBvKc;
  g.pad = p->pad;
  g.tiles_y = div_up(g.h_out, kBvTileY);
  g.tiles_x = div_up(g.w_out, kBvTileX);
  g.groups = p->groups; g.cgroups = p->cgroups; g.up = p->up;
  g.out_h = g.h_out * p->up; g.out_w = g.w_out * p->up;
  g.out_channels = p->out_channels; g.out_c0 = p->out_c0;
  static std::atomic<int> launch_seq{0};
  g.seq = launch_seq.fetch_add(1, std::memory_order_relaxed);
  // Channel-stationary schedule where it pays (measured, scratch/bev_cs_probe.py): with several tiles per CTA it is ~10 %
  // faster (fewer operand reads from shared memory, half the MMA instructions); with one tile per CTA (SECOND: 143 tiles)
  // both schedules take the same time inside the product graph, and the pixel-stationary epilogue is the shorter one.
  const int variant = bev_variant();
  const int n_tiles_all = g.batch * g.tiles_y * g.tiles_x * g.groups;
  if ((variant == 1 || (variant == 2 && n_tiles_all >= 2 * kNumSMs)) && p->ksize == 3 && p->stride == 1 && p->c_out == kBwCout &&
      p->c_in % kBvKc == 0 && p->up == 1 && p->out_channels % 4 == 0)
    return launch_bev_cs(p, g, stream);
#define D3B_BEV_CASE(KS, ST)                                                   \
  if (p->ksize == KS && p->stride == ST) {                                     \
    switch (p->c_out) {                                                        \
      case 32: return launch_bev<KS, ST, 32>(p, g, stream);                    \
      case 64: return launch_bev<KS, ST, 64>(p, g, stream);                    \
      case 128: return launch_bev<KS, ST, 128>(p, g, stream);                  \
      default: break;                                                          \
    }                                                                          \
  }
  D3B_BEV_CASE(3, 1)
  D3B_BEV_CASE(3, 2)
  D3B_BEV_CASE(1, 1)
#undef D3B_BEV_CASE
  set_error("d3b_bev_conv16: C_out per group %d not in {32, 64, 128}", p->c_out);
  return D3B_ERR_UNSUPPORTED;
}

#ifdef D3B_SOFT_TIMEOUT
extern "C" int d3b_debug_fault_bevconv16(unsigned int* host8) {
  cudaError_t e = cudaMemcpyFromSymbol(host8, d3b::g_d3b_fault, 32);
  unsigned int zeros[8] = {0};
  if (e == cudaSuccess) e = cudaMemcpyToSymbol(d3b::g_d3b_fault, zeros, 32);
  return (int)e;
}
extern "C" int d3b_debug_cta_ns_bevconv16(unsigned long long* host4096) {
  return (int)cudaMemcpyFromSymbol(host4096, d3b::g_d3b_cta_ns, sizeof(unsigned long long) * 4096);
}
extern "C" int d3b_debug_cta_clk_bevconv16(long long* host4096) {
  return (int)cudaMemcpyFromSymbol(host4096, d3b::g_d3b_cta_clk, sizeof(long long) * 4096);
}
extern "C" int d3b_debug_trace_bevconv16(long long* host, int clear) {
  cudaError_t e = cudaMemcpyFromSymbol(host, d3b::g_d3b_trace, sizeof(long long) * 16 * 512);
  if (e == cudaSuccess && clear) {
    static long long zeros[16 * 512];
    e = cudaMemcpyToSymbol(d3b::g_d3b_trace, zeros, sizeof(zeros));
  }
  return (int)e;
}
#endif
