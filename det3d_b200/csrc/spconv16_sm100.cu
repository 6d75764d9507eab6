// Output-stationary sparse convolution on tcgen05 with split-f16 operands ("FP16x3"), sm_100a.
//
//   out[o,:] = act((sum_k in[nbr[k][o],:] . W[k] + bias) * scale + shift + residual[o,:])
//
// Reference: spconv v1.x indice_conv (per offset: gather -> fp32 torch.mm -> scatter-add) followed by
// BatchNorm1d(eval) / ReLU / residual, call sites det3d/models/backbones/scn.py:73-89,106-157,323-355.
//
// Why this form (round 2).  The pair-based kernel (sparse_conv_sm100.cu, D3B_ALGO_TC_PAIRS) sums the offsets' partial
// products with fp32 atomics: summation order -- hence the last bits -- changed run to run, its 14 launches needed 8
// buffer-clearing launches and a deferred epilogue, and at lidar densities (~100-160 work items per layer) it was bound
// by the RED issue rate and two dependent L2 round trips per item.  Here one CTA owns 128 output rows and walks the
// kernel offsets present in the tile (tile_mask); an (offset, 64-channel slice) is one pipeline slot: gather warps copy
// the 128 input rows (or zeros where the offset has no neighbour) with 16-byte cp.async straight into a K-major,
// 128B-swizzled tile, one warp issues the MMAs, dedicated accumulator warps keep the running sums in registers, and the
// fused bias/BN/residual/ReLU epilogue writes each output row once.  No atomics, fixed summation order: bit-identical
// results run to run, and a row's bits depend on its own neighbourhood only (not on the rows sharing its tile).
//
// fp32-equivalent accuracy on the f16 pipe (twice the tf32 rate, half the operand bytes).  Activations live in HBM as
// two f16 planes, hi = f16(x) and lo = f16(x - hi) (22 significant bits, written once by the producing layer's
// epilogue); weights are split the same way at load time after an exact power-of-two scaling that keeps their lo
// parts out of the f16 subnormal range.  D += A_hi.B_hi + A_hi.B_lo + A_lo.B_hi with fp32 accumulation; the dropped
// lo.lo term is 2^-22 relative.  |x| >= 65504 cannot be represented: the epilogue raises a device flag instead of
// silently saturating (the host checks it with the detections).
//
// Accumulation.  The tensor core adds every MMA's partial sum into the fp32 TMEM accumulator with truncation (round
// toward zero): measured on this path, the relative error of a layer grew linearly with the number of MMAs chained into
// one accumulator, about -2^-26 per MMA (27 offsets x 8..24 MMAs -> 2e-5 .. 7e-5 over the encoder), a systematic shrink
// that 20 layers turn into > 1e-4.  So (a) the chain is bounded: a TMEM buffer collects kFlush = 3 slots of the FULL slot
// enumeration (fixed ranges, whichever of them the tile mask activates), then the accumulator warps add it into fp32
// registers with round-to-nearest while the next group fills the other buffer -- the summation structure of the
// reference's per-offset GEMM + scatter-add, in a fixed order; and (b) the mean of the truncation is undone: a row's
// partial sum is scaled by 1 + n * 2^-26, n = the MMAs that contributed to THAT row (absent neighbours add exact zeros).
//
// Roles: 4 or 8 accumulator warps (first, so they get the registers), one gather group of 2 warps per pipeline stage
// (slot j -> group j % stages; 4 stages, 3 at C_out = 128; pure cp.async issue), 1 MMA warp (also owns TMEM); persistent
// grid <= 148 CTAs.  Per k-step two MMAs: A_hi x [B_hi | B_lo] (N = 2 C_out) and A_lo x B_hi (N = C_out).  C_in = 16 / 32
// layers pack 4 / 2 kernel offsets into one slot (os16_pack).  The tile's rulebook rows arrive by cp.async.bulk; the
// kernel is launched with programmatic stream serialisation (prologue overlaps the previous layer's tail).
// Algorithmic bytes per layer: N_in*C_in*4 + N_out*C_out*4 + P*8 + K*C_in*C_out*4.
#include "umma.cuh"

namespace d3b {

constexpr int kOsTileM = 128;
constexpr int kOsKc = 64;                                  // channels per stage = one 128-byte swizzle row of f16
constexpr int kOsABytes = kOsTileM * 128;                  // one A plane tile (hi or lo)

template <int COUT>
struct OsCfg {
  static constexpr int kBBytes = 2 * COUT * 128;           // [B_hi rows | B_lo rows]
  static constexpr int kStageBytes = 2 * kOsABytes + kBBytes;
  static constexpr int kStages = COUT >= 128 ? 3 : 4;
  // One gather group per pipeline stage: group g only ever fills stage g.  (A group waits for "stage free" on the
  // PARITY of the stage's empty barrier; that is unambiguous only if the group itself has seen the previous use of
  // the stage go by -- with more groups than stages a group would meet a stage it last touched two uses ago, read a
  // stale parity as "free" and overwrite live operands.)
  static constexpr int kGroups = kStages;
  // Accumulator warps: every TMEM lane quadrant is served by kEpiWarps / 4 warps, each owning kCols columns of the
  // output row its lane stands for (the running fp32 sums live in registers, see the kernel comment).
  static constexpr int kEpiWarps = COUT >= 64 ? 8 : 4;
  static constexpr int kCols = COUT / (kEpiWarps / 4);     // columns per accumulator thread: 16, 32, 32, 64
  static constexpr int kGatherWarp0 = kEpiWarps;           // warps [kEpiWarps, kEpiWarps + 2 kGroups): gather groups of 2 warps
  static constexpr int kGroupThreads = 64;                 // (few threads: the register file goes to the accumulators)
  static constexpr int kMmaWarp = kEpiWarps + 2 * kGroups; // TMEM alloc + MMA issue
  static constexpr int kThreads = 32 * (kMmaWarp + 1);     // 416 (C_out 16/32), 544 (64), 480 (128)
  // Accumulators.  Per k-step TWO MMAs: A_hi x [B_hi | B_lo] (N = 2 C_out: the hi.hi and hi.lo products side by side)
  // and A_lo x B_hi (N = C_out, onto the hi.hi columns) -- a single thread issues only ~one tcgen05.mma per 65..90
  // cycles (clock-stamp trace), three N = C_out MMAs per k-step made the issue rate the bottleneck.
  static constexpr int kBufCols = 2 * COUT;                // [hi.hi + lo.hi | hi.lo]
  static constexpr int kTmemCols = 2 * kBufCols;           // two buffers: group g+1 accumulates while group g is drained
  // An accumulator buffer lives for kFlush slots (offsets), then the accumulator warps add it into fp32 registers: the
  // chain of truncating tensor-core accumulations stays <= kFlush * 8 MMAs, and the TMEM read (the slow direction)
  // is amortised over kFlush slots.
  static constexpr int kFlush = 3;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/ + 32 * kOsTileM * 4;
};

__device__ __forceinline__ uint32_t pack_half2(__half a, __half b) {
  return (uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16);
}

// Fused epilogue on 16 consecutive output channels of one row; returns true if a value left the f16 range.
struct OsEpi {
  const float* bias;
  const float* scale;
  const float* shift;
  const __half* res_hi;
  const __half* res_lo;
  float acc_scale;
  int relu;
  int seq;                        // launch counter of this translation unit (development traces only)
};

__device__ __forceinline__ bool epilogue16(float (&v)[16], const OsEpi& e, size_t row_off, int col, __half* out_hi,
                                           __half* out_lo, float* out_f32) {
#pragma unroll
  for (int q = 0; q < 16; ++q) v[q] *= e.acc_scale;
  if (e.bias) {
#pragma unroll
    for (int q = 0; q < 16; q += 4) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(e.bias + col + q));
      v[q] += b.x; v[q + 1] += b.y; v[q + 2] += b.z; v[q + 3] += b.w;
    }
  }
  if (e.scale) {
#pragma unroll
    for (int q = 0; q < 16; q += 4) {
      const float4 s = __ldg(reinterpret_cast<const float4*>(e.scale + col + q));
      const float4 t = __ldg(reinterpret_cast<const float4*>(e.shift + col + q));
      v[q] = fmaf(v[q], s.x, t.x); v[q + 1] = fmaf(v[q + 1], s.y, t.y);
      v[q + 2] = fmaf(v[q + 2], s.z, t.z); v[q + 3] = fmaf(v[q + 3], s.w, t.w);
    }
  }
  if (e.res_hi) {
#pragma unroll
    for (int q = 0; q < 16; q += 8) {
      const uint4 h = __ldg(reinterpret_cast<const uint4*>(e.res_hi + row_off + col + q));
      const uint4 l = __ldg(reinterpret_cast<const uint4*>(e.res_lo + row_off + col + q));
      const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 fh = __half22float2(*reinterpret_cast<const __half2*>(&hw[j]));
        const float2 fl = __half22float2(*reinterpret_cast<const __half2*>(&lw[j]));
        v[q + 2 * j] += fh.x + fl.x;          // hi + lo is exact in fp32 (22 bits)
        v[q + 2 * j + 1] += fh.y + fl.y;
      }
    }
  }
  bool ovf = false;
  if (e.relu) {
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
      hi[q >> 1] = pack_half2(h0, h1);
      lo[q >> 1] = pack_half2(l0, l1);
      ovf |= !(fabsf(v[q]) < 65504.f) | !(fabsf(v[q + 1]) < 65504.f);
    }
    uint4* ph = reinterpret_cast<uint4*>(out_hi + row_off + col);
    uint4* pl = reinterpret_cast<uint4*>(out_lo + row_off + col);
    ph[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    ph[1] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
    pl[0] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    pl[1] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
  }
  if (out_f32) {
    float4* pf = reinterpret_cast<float4*>(out_f32 + row_off + col);
#pragma unroll
    for (int q = 0; q < 16; q += 4) pf[q >> 2] = make_float4(v[q], v[q + 1], v[q + 2], v[q + 3]);
  }
  return ovf;
}

// Offset packing.  A pipeline slot always moves 128 rows x 128 bytes per plane and costs about the same whatever part
// of it is live, so layers with C_in = 16 / 32 put `pack` = 4 / 2 kernel offsets side by side along K: slot p holds the
// rows of offsets p*pack .. p*pack + pack - 1 (16-byte chunks [sub * 8/pack, (sub+1) * 8/pack) of every row come from
// offset p*pack + sub), the weight slice is the matching K-concatenation (zero for offsets >= k_vol), and a 3x3x3 layer
// runs 14 / 7 slots per tile instead of 27.  A slot is skipped when none of its offsets occurs in the tile.
__host__ __device__ inline int os16_pack(int c_in, int k_vol) {
  return k_vol > 1 && (c_in == 16 || c_in == 32) ? kOsKc / c_in : 1;
}
__device__ __forceinline__ unsigned int os16_slot_mask(unsigned int offset_mask, int pack) {
  if (pack == 1) return offset_mask;
  unsigned int r = 0u;
  const unsigned int grp = (1u << pack) - 1u;
  for (int p = 0; p * pack < 32; ++p) r |= ((offset_mask >> (p * pack)) & grp) ? (1u << p) : 0u;
  return r;
}

template <int COUT>
__global__ void __launch_bounds__(OsCfg<COUT>::kThreads, 1)
spconv_os16_kernel(const __half* __restrict__ in_hi, const __half* __restrict__ in_lo, const int* __restrict__ nbr,
                   const unsigned int* __restrict__ tile_mask, const int* __restrict__ n_out_p, int out_cap, int c_in,
                   int n_kb, int pack, int k_vol, const __half* __restrict__ packed, OsEpi epi, __half* __restrict__ out_hi,
                   __half* __restrict__ out_lo, float* __restrict__ out_f32, int* __restrict__ overflow) {
  using Cfg = OsCfg<COUT>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + Cfg::kStages * Cfg::kStageBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::kStages + s); };
  auto acc_full = [&](uint32_t b) { return bar_base + 8u * (2 * Cfg::kStages + b); };
  auto acc_empty = [&](uint32_t b) { return bar_base + 8u * (2 * Cfg::kStages + 2 + b); };
  const uint32_t nbr_bar = bar_base + 8u * (2 * Cfg::kStages + 4);                 // neighbour rows of the tile have landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_gen + Cfg::kStages * Cfg::kStageBytes + 8 * (2 * Cfg::kStages + 5));
  int* koff_s = reinterpret_cast<int*>(smem_gen + Cfg::kStages * Cfg::kStageBytes + 128);   // [32] active offsets
  int* nbr_s = reinterpret_cast<int*>(smem_gen + Cfg::kStages * Cfg::kStageBytes + 256);    // [32][128] neighbour rows

  D3B_CTA_MARK(0, epi.seq);
  pdl_launch_dependents();           // the next kernel of the stream may start its prologue behind this one's tail
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_out = min(*n_out_p, out_cap);
  const int n_tiles = (n_out + kOsTileM - 1) / kOsTileM;
  const int c_eff = c_in * pack;                 // K extent of the slots (64 when offsets are packed)
  const int n_full = ((k_vol + pack - 1) / pack) * n_kb;     // slots of a tile in which every offset occurs

  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(full_bar(s), Cfg::kGroupThreads + 1);   // the gather threads of one group + the expect_tx arrive
      mbar_init(empty_bar(s), 1);        // tcgen05.commit
    }
    for (uint32_t b = 0; b < 2; ++b) {
      mbar_init(acc_full(b), 1);                       // tcgen05.commit of a slot's MMAs
      mbar_init(acc_empty(b), 32 * Cfg::kEpiWarps);    // every accumulator warp has read the slot's partial sums
    }
    mbar_init(nbr_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == Cfg::kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)Cfg::kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;

  if (warp < Cfg::kEpiWarps) {
    // ===================== accumulator warps =====================
    pdl_wait_prior_grid();           // (residual planes / output buffers belong to earlier kernels)
    // Per slot: TMEM partial sums -> fp32 registers (round-to-nearest adds).  At the end of the tile: fused
    // bias / BN / residual / ReLU, split into f16 planes, one store per output row.
    const int quad = warp & 3;                       // TMEM lane quadrant this warp may read
    const int col0 = (warp >> 2) * Cfg::kCols;       // first output column owned by this thread
    bool ovf = false;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const unsigned int smask = os16_slot_mask(tile_mask[tile], pack);
      const int o = tile * kOsTileM + quad * 32 + lane;
      float acc[Cfg::kCols];
#pragma unroll
      for (int q = 0; q < Cfg::kCols; ++q) acc[q] = 0.f;
      // which kernel offsets reach THIS row: an absent neighbour contributes exact zeros (nothing is truncated), so the
      // truncation correction of a group counts the row's own live k-steps -- with the fixed grouping below, a row's bits
      // depend on its own neighbourhood only, never on which other rows share its tile (batch composition)
      unsigned int live = 0u;
      if (o < n_out) {
#pragma unroll 9
        for (int k = 0; k < k_vol; ++k) live |= (__ldg(nbr + (size_t)k * out_cap + o) >= 0 ? 1u : 0u) << k;
      }
      for (int f0 = 0; f0 < n_full; f0 += Cfg::kFlush) {
        int n_active = 0, n_ks_row = 0;                 // slots of the group present in the tile; this row's live k-steps
        for (int f = f0; f < min(f0 + Cfg::kFlush, n_full); ++f) {
          const int sp = n_kb == 1 ? f : f / n_kb, kb = f - sp * n_kb;      // (no division on the usual path)
          if (!((smask >> sp) & 1u)) continue;
          ++n_active;
          if (pack == 1)
            n_ks_row += ((live >> sp) & 1u) ? min(kOsKc / 16, (c_eff - kb * kOsKc + 15) / 16) : 0;
          else
            n_ks_row += __popc((live >> (sp * pack)) & ((1u << pack) - 1u)) * (kOsKc / 16 / pack);
        }
        if (n_active == 0) continue;
        const uint32_t buf = it & 1u;
        if (warp == 0 && lane == 0) D3B_STAMP(8, it);
        D3B_WAIT(acc_full(buf), (it >> 1) & 1u, 4);
        if (warp == 0 && lane == 0) D3B_STAMP(9, it);
        tc_fence_after();
        const uint32_t t0 = tmem_d + buf * Cfg::kBufCols + ((uint32_t)(quad * 32) << 16) + col0;
        const float f_hh = 1.f + 2.f * (float)n_ks_row * kTruncLossPerMma;     // two MMAs per k-step land on these columns
        const float f_hl = 1.f + (float)n_ks_row * kTruncLossPerMma;
#pragma unroll
        for (int c0 = 0; c0 < Cfg::kCols; c0 += 16) {
          uint32_t r0[16], r1[16];
          tc_ld16_nowait(t0 + c0, r0);                 // hi.hi + lo.hi
          tc_ld16_nowait(t0 + COUT + c0, r1);          // hi.lo
          tc_ld_wait();
#pragma unroll
          for (int q = 0; q < 16; ++q)
            acc[c0 + q] += fmaf(__uint_as_float(r0[q]), f_hh, __uint_as_float(r1[q]) * f_hl);
        }
        tc_fence_before();
        mbar_arrive(acc_empty(buf));               // the MMA thread may overwrite this buffer
        if (warp == 0 && lane == 0) D3B_STAMP(10, it);
        ++it;
      }
      if (o < n_out) {
#pragma unroll
        for (int c0 = 0; c0 < Cfg::kCols; c0 += 16) {
          float v[16];
#pragma unroll
          for (int q = 0; q < 16; ++q) v[q] = acc[c0 + q];
          ovf |= epilogue16(v, epi, (size_t)o * COUT, col0 + c0, out_hi, out_lo, out_f32);
        }
      }
    }
    if (ovf && overflow) atomicOr(overflow, 1);
  } else if (warp < Cfg::kMmaWarp) {
    // ===================== gather producers =====================
    const int gw = warp - Cfg::kGatherWarp0;
    const int group = gw >> 1, wq = gw & 1;                     // two warps per group, 64 rows each
    const int g = lane >> 3, c = lane & 7;
    const bool issues_tma = (wq == 0 && lane == 0);
    const int ptid = threadIdx.x - 32 * Cfg::kGatherWarp0;      // 0 .. 64 * kGroups - 1
    const int c_in_pad = (c_in + 15) & ~15;
    const int cpo = 8 / pack;                                   // 16-byte chunks per packed offset
    const int sub = c / cpo;                                    // which of the slot's offsets feeds this thread's chunk
    const int c_src = c - sub * cpo;                            // ... and which chunk of that offset's source row
    uint32_t it0 = 0;       // pipeline slots consumed by earlier tiles (same sequence in every role)
    uint32_t tile_count = 0;                       // tiles with work so far (phase of nbr_bar)
    const bool bulk_nbr = (out_cap & 3) == 0;      // 16-byte aligned rows: cp.async.bulk can fetch them
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int row0 = tile * kOsTileM;
      const unsigned int mask = os16_slot_mask(tile_mask[tile], pack);
      const int n_off = __popc(mask);                 // slots' worth of offsets (groups of `pack`) present in the tile
      const int n_slots = n_off * n_kb;

      // stage nbr[k][row0 .. row0+127] for the active offsets (one global round trip per tile)
      asm volatile("bar.sync 1, %0;" ::"r"(Cfg::kGroupThreads * Cfg::kGroups) : "memory");   // previous tile's readers are done
      if (ptid < 32) {                                    // n-th active offset of the tile
        unsigned int m = mask;
        for (int t = ptid; t > 0; --t) m &= m - 1;
        const int k = __ffs(m) - 1;
        if (ptid < n_off) koff_s[ptid] = k;
        __syncwarp();
        // neighbour rows of the tile by bulk copy (one 512-byte copy per offset of an active slot, no thread touches
        // them): entry e = (slot e / pack, member e % pack) -> nbr_s[e][0..127]
        const int rows = min(kOsTileM, out_cap - row0);
        const int e_slot = ptid / pack;
        const int k_src = ptid < n_off * pack ? koff_s[e_slot] * pack + (ptid - e_slot * pack) : k_vol;
        const bool valid = k_src < k_vol;             // (the last group of a 27-offset kernel has phantom members)
        const unsigned int vm = __ballot_sync(0xffffffffu, valid);
        if (bulk_nbr && vm != 0u && ptid == 0) mbar_arrive_expect_tx(nbr_bar, (uint32_t)(__popc(vm) * rows * 4));
        __syncwarp();
        if (bulk_nbr && valid)
          tma_bulk_g2s(smem_u32(nbr_s + ptid * kOsTileM), nbr + (size_t)k_src * out_cap + row0, (uint32_t)(rows * 4), nbr_bar);
      }
      if (bulk_nbr) {
        if (n_off > 0) D3B_WAIT(nbr_bar, tile_count & 1u, 5);
      } else {
        asm volatile("bar.sync 1, %0;" ::"r"(Cfg::kGroupThreads * Cfg::kGroups) : "memory");
#pragma unroll 8
        for (int idx = ptid; idx < n_off * pack * kOsTileM; idx += Cfg::kGroupThreads * Cfg::kGroups) {   // independent loads
          const int r = idx & 127, e = idx >> 7;
          const int k_src = koff_s[e / pack] * pack + e % pack;
          nbr_s[idx] = k_src < k_vol ? __ldg(nbr + (size_t)k_src * out_cap + min(row0 + r, out_cap - 1)) : -1;
        }
      }
      asm volatile("bar.sync 1, %0;" ::"r"(Cfg::kGroupThreads * Cfg::kGroups) : "memory");
      if (n_off > 0) ++tile_count;
      if (tile == (int)blockIdx.x) pdl_wait_prior_grid();     // rulebook rows staged; the activations need the previous layer done

      for (int j = (int)((group + Cfg::kGroups - (it0 % Cfg::kGroups)) % Cfg::kGroups); j < n_slots; j += Cfg::kGroups) {
        const int n = j / n_kb, kb = j - n * n_kb;
        const int ch = pack > 1 ? c_src * 8 : kb * kOsKc + c * 8;     // this thread's 8 channels (16 bytes) of the source row
        const uint32_t it = it0 + (uint32_t)j;
        const int s = it % Cfg::kStages;
        const uint32_t ph = (it / Cfg::kStages) & 1u;
        if (issues_tma) D3B_STAMP(0, it);
        D3B_WAIT(empty_bar(s), ph ^ 1u, 1);
        if (issues_tma) D3B_STAMP(1, it);
        const uint32_t stage = smem_base + s * Cfg::kStageBytes;
        if (issues_tma) {
          mbar_arrive_expect_tx(full_bar(s), Cfg::kBBytes);
          tma_bulk_g2s(stage + 2 * kOsABytes, packed + ((size_t)koff_s[n] * n_kb + kb) * (Cfg::kBBytes / 2), Cfg::kBBytes,
                       full_bar(s));
        }
        if (pack > 1 || ch < c_in_pad) {
          // this thread: 16 consecutive rows (indices fetched as four 16-byte shared-memory loads), one 16-byte chunk
          const int row_base = wq * 64 + g * 16;
          const int4* idx4 = reinterpret_cast<const int4*>(nbr_s + (n * pack + sub) * kOsTileM + row_base);
          const bool col_live = pack > 1 ? koff_s[n] * pack + sub < k_vol : ch < c_in;
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            const int4 iv = idx4[q4];
            const int srcs[4] = {iv.x, iv.y, iv.z, iv.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int row = row_base + q4 * 4 + u;
              const bool live = srcs[u] >= 0 && col_live && row0 + row < n_out;    // (rows past n_out hold stale indices)
              const size_t off = live ? (size_t)srcs[u] * c_in + ch : 0;
              const uint32_t dst = stage + sw128_offset(row, c);
              cp_async16(dst, in_hi + off, live ? 16u : 0u);
              cp_async16(dst + kOsABytes, in_lo + off, live ? 16u : 0u);
            }
          }
        }
        if (issues_tma) D3B_STAMP(2, it);
        cp_async_wait_all();
        if (issues_tma) D3B_STAMP(3, it);
        fence_proxy_async();      // generic-proxy writes -> visible to the tensor core (async proxy)
        mbar_arrive(full_bar(s));
      }
      it0 += (uint32_t)n_slots;
    }
  } else {
    // ===================== MMA issuer (one elected lane) =====================
    // Every slot (offset, 64-channel slice) is accumulated from zero in its own TMEM buffer, its <= 12 MMAs spread over
    // group's buffer (<= 24 chained truncating accumulations); everything beyond is summed by the accumulator warps.
    constexpr uint32_t idesc2 = umma_idesc_f16(kOsTileM, 2 * COUT);   // A_hi x [B_hi | B_lo]
    constexpr uint32_t idesc1 = umma_idesc_f16(kOsTileM, COUT);       // A_lo x B_hi
    uint32_t it = 0, git = 0;          // slot counter (operand stages), group counter (accumulator buffers)
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const unsigned int smask = os16_slot_mask(tile_mask[tile], pack);
      // accumulator groups are FIXED ranges of the full slot enumeration f = slot_offset * n_kb + kb (not runs of the
      // slots that happen to be active in this tile): which partial sums share a truncating TMEM chain must not depend
      // on the other rows of the tile
      for (int f0 = 0; f0 < n_full; f0 += Cfg::kFlush) {
        // (n_kb == 1 unless C_in > 64: no division on the usual path -- this warp's instruction stream is the critical one)
        unsigned int members = 0u;                      // bit i: slot f0 + i is active in this tile
        for (int i = 0; i < Cfg::kFlush && f0 + i < n_full; ++i)
          members |= ((smask >> (n_kb == 1 ? f0 + i : (f0 + i) / n_kb)) & 1u) << i;
        if (members == 0u) continue;
        const uint32_t buf = git & 1u;
        if (lane == 0) D3B_STAMP(4, git);
        D3B_WAIT(acc_empty(buf), ((git >> 1) & 1u) ^ 1u, 2);
        if (lane == 0) D3B_STAMP(5, git);
        const uint32_t d_addr = tmem_d + buf * Cfg::kBufCols;
        bool first = true;
        for (int i = 0; i < Cfg::kFlush; ++i) {
          if (!((members >> i) & 1u)) continue;
          const int kb = n_kb == 1 ? 0 : (f0 + i) % n_kb;
          const int s = it % Cfg::kStages;
          const uint32_t ph = (it / Cfg::kStages) & 1u;
          const int n_ks = min(kOsKc / 16, (c_eff - kb * kOsKc + 15) / 16);
          D3B_WAIT(full_bar(s), ph, 3);
          if (lane == 0) D3B_STAMP(6, it);
          tc_fence_after();
          {
            // warp-uniform instruction stream (descriptor arithmetic stays in uniform registers); lane 0 issues
            const uint32_t issue = lane == 0 ? 1u : 0u;
            const uint32_t a_hi = smem_base + s * Cfg::kStageBytes;
            const uint64_t da_hi = umma_desc_sw128(a_hi), da_lo = umma_desc_sw128(a_hi + kOsABytes);
            const uint64_t db = umma_desc_sw128(a_hi + 2 * kOsABytes);
#pragma unroll
            for (int ks = 0; ks < kOsKc / 16; ++ks) {
              if (ks < n_ks) {
                const uint64_t adv = (uint64_t)(ks * 2);   // 16 f16 = 32 bytes along K = 2 descriptor address units
                tc_mma_f16_if(issue, d_addr, da_hi + adv, db + adv, idesc2, (!first || ks > 0) ? 1u : 0u);
                tc_mma_f16_if(issue, d_addr, da_lo + adv, db + adv, idesc1, 1u);
              }
            }
            tc_commit_if(issue, empty_bar(s));      // frees the operand stage when these MMAs have read it
            if (lane == 0) D3B_STAMP(7, it);
          }
          first = false;
          ++it;
        }
        tc_commit_if(lane == 0 ? 1u : 0u, acc_full(buf));     // hands the group's partial sums to the accumulator warps
        ++git;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  D3B_CTA_MARK(1, epi.seq);
  if (warp == Cfg::kMmaWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"((uint32_t)Cfg::kTmemCols)
                 : "memory");
  }
}

// ---- first layer: fp32 rows with a handful of channels (C_in <= 16: the voxel mean, 4 or 5 features) -----------------
// 2*27*C_in*C_out flops per row -- nothing for the tensor cores.  fp32 FFMA, output-stationary, same epilogue and
// output format as the tcgen05 kernel.  Four lanes share a (row, 16-column part), each taking every fourth offset (their
// neighbour-index loads are independent and in flight together: the kernel is pure latency), combined by a fixed shuffle tree.
template <int COUT>
__global__ void __launch_bounds__(256)
spconv_first16_kernel(const float* __restrict__ feat_in, const int* __restrict__ nbr, const int* __restrict__ n_out_p,
                      int out_cap, int c_in, int k_vol, const float* __restrict__ weight, OsEpi epi,
                      __half* __restrict__ out_hi, __half* __restrict__ out_lo, float* __restrict__ out_f32,
                      int* __restrict__ overflow) {
  extern __shared__ float w_s[];                 // [k_vol][c_in][COUT]
  for (int i = threadIdx.x; i < k_vol * c_in * COUT; i += blockDim.x) w_s[i] = weight[i];
  __syncthreads();
  constexpr int kParts = COUT / 16;              // 16-column parts of a row
  constexpr int kSub = 4;                        // lanes per (row, part): lane `sub` takes the offsets k = sub, sub + 4, ...
  constexpr int kMaxK = 8;                       // ceil(32 / 4) offsets per lane at most
  const int n_out = min(*n_out_p, out_cap);
  const long long items = (long long)n_out * kParts;
  const long long items_pad = (items + 7) / 8 * 8;          // whole warps (8 items each) run the shuffles together
  bool ovf = false;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < items_pad * kSub; e += (long long)gridDim.x * blockDim.x) {
    const long long item = e / kSub;
    const int sub = (int)(e % kSub);
    const bool on = item < items;
    const int o = on ? (int)(item / kParts) : 0, col = (int)(item % kParts) * 16;
    int src[kMaxK];
#pragma unroll
    for (int t = 0; t < kMaxK; ++t) {            // the neighbour indices first: independent loads, all in flight together
      const int k = sub + t * kSub;
      src[t] = (on && k < k_vol) ? __ldg(nbr + (size_t)k * out_cap + o) : -1;
    }
    float acc[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) acc[q] = 0.f;
#pragma unroll
    for (int t = 0; t < kMaxK; ++t) {
      if (src[t] < 0) continue;
      const int k = sub + t * kSub;
      for (int ci = 0; ci < c_in; ++ci) {
        const float a = __ldg(feat_in + (size_t)src[t] * c_in + ci);
        const float* w = w_s + ((size_t)k * c_in + ci) * COUT + col;
#pragma unroll
        for (int q = 0; q < 16; ++q) acc[q] = fmaf(a, w[q], acc[q]);
      }
    }
#pragma unroll
    for (int q = 0; q < 16; ++q) {               // fixed-order tree over the four offset classes
      acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], 1);
      acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], 2);
    }
    if (on && sub == 0) ovf |= epilogue16(acc, epi, (size_t)o * COUT, col, out_hi, out_lo, out_f32);
  }
  if (ovf && overflow) atomicOr(overflow, 1);
}

// ---- f16 weight image ------------------------------------------------------------------------------------------------
// packed[slot][kb][part][n][swizzled 64 halves], part 0 = hi, 1 = lo of w * 2^w_exp; zero beyond c_in.  slot = kernel
// offset, or with offset packing (C_in 16 / 32, see os16_pack) a group of `pack` offsets side by side along K.
__global__ void __launch_bounds__(256)
pack_weight16_kernel(const float* __restrict__ w, int c_in, int c_out, int k_vol, int n_slots_k, int n_kb, int pack,
                     float w_mul, __half* __restrict__ packed) {
  const long long total = (long long)n_slots_k * n_kb * 2 * c_out * kOsKc;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    long long t = e;
    const int cc = (int)(t % kOsKc); t /= kOsKc;
    const int n = (int)(t % c_out); t /= c_out;
    const int part = (int)(t % 2); t /= 2;
    const int kb = (int)(t % n_kb); t /= n_kb;
    int k = (int)t;                                  // slot
    int ci = kb * kOsKc + cc;
    if (pack > 1) { k = k * pack + cc / c_in; ci = cc % c_in; }
    const float x = (ci < c_in && k < k_vol) ? w[((size_t)k * c_in + ci) * c_out + n] * w_mul : 0.0f;
    k = (int)t;
    __half hi, lo;
    split_f16(x, hi, lo);
    const size_t tile = (((size_t)k * n_kb + kb) * 2 + part) * (size_t)(c_out * kOsKc);
    const uint32_t off = sw128_offset(n, cc >> 3) + (cc & 7) * 2;     // bytes
    packed[tile + off / 2] = part == 0 ? hi : lo;
  }
}

// ---- plane <-> fp32 conversions (API boundary, tests) ------------------------------------------------------------------
__global__ void __launch_bounds__(256)
split16_kernel(const float* __restrict__ x, long long n, __half* __restrict__ hi, __half* __restrict__ lo,
               int* __restrict__ overflow) {
  bool ovf = false;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const float v = x[e];
    __half h, l;
    split_f16(v, h, l);
    hi[e] = h;
    lo[e] = l;
    ovf |= !(fabsf(v) < 65504.f);
  }
  if (ovf && overflow) atomicOr(overflow, 1);
}

__global__ void __launch_bounds__(256)
merge16_kernel(const __half* __restrict__ hi, const __half* __restrict__ lo, long long n, float* __restrict__ x) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x)
    x[e] = __half2float(hi[e]) + __half2float(lo[e]);
}

// rows (f16 planes or fp32) -> channels-last BEV planes [B*H*W, C*D] (pre-zeroed), channel = c*D + z: the values of
// `dense.view(B, C*D, H, W)` (scn.py:192-195) in NHWC order.
__global__ void __launch_bounds__(256)
sparse_to_bev16_kernel(const __half* __restrict__ in_hi, const __half* __restrict__ in_lo, const float* __restrict__ in_f32,
                       const int* __restrict__ coors, const int* __restrict__ n_rows, int row_cap, int C, int D, int H,
                       int W, int B, __half* __restrict__ out_hi, __half* __restrict__ out_lo) {
  const int n = min(*n_rows, row_cap);
  const long long total = (long long)n * C;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(e / C), c = (int)(e - (long long)r * C);
    const int4 q = *reinterpret_cast<const int4*>(coors + (size_t)r * 4);
    if ((unsigned)q.x >= (unsigned)B || (unsigned)q.y >= (unsigned)D || (unsigned)q.z >= (unsigned)H ||
        (unsigned)q.w >= (unsigned)W)
      continue;
    const size_t dst = (((size_t)q.x * H + q.z) * W + q.w) * ((size_t)C * D) + (size_t)c * D + q.y;
    if (in_f32) {
      __half h, l;
      split_f16(in_f32[e], h, l);
      out_hi[dst] = h;
      out_lo[dst] = l;
    } else {
      out_hi[dst] = in_hi[e];
      out_lo[dst] = in_lo[e];
    }
  }
}

static bool os16_shape_ok(int c_in, int c_out) {
  const bool cin_ok = c_in >= 8 && c_in <= 512 && c_in % 8 == 0;    // rows are gathered in 16-byte chunks
  const bool cout_ok = c_out == 16 || c_out == 32 || c_out == 64 || c_out == 128;
  return cin_ok && cout_ok;
}

static OsEpi epi_of(const d3b_conv16_params* p) {
  OsEpi e;
  e.bias = p->bias; e.scale = p->scale; e.shift = p->shift;
  e.res_hi = (const __half*)p->residual_hi; e.res_lo = (const __half*)p->residual_lo;
  e.acc_scale = p->acc_scale; e.relu = p->relu;
  static std::atomic<int> launch_seq{0};
  e.seq = launch_seq.fetch_add(1, std::memory_order_relaxed);
  return e;
}

template <int COUT>
static int launch_os16(const d3b_conv16_params* p, const int32_t* nbr, const uint32_t* tile_mask, const int32_t* n_out,
                       int32_t out_cap, cudaStream_t stream) {
  using Cfg = OsCfg<COUT>;
  static SmemOptIn optin;
  D3B_CUDA(ensure_dynamic_smem(spconv_os16_kernel<COUT>, Cfg::kSmemBytes, optin));
  const int n_tiles = div_up(out_cap, kOsTileM);
  const int grid = n_tiles < kNumSMs ? (n_tiles > 0 ? n_tiles : 1) : kNumSMs;
  const int pack = os16_pack(p->c_in, p->k_vol);
  const int n_kb = pack > 1 ? 1 : (p->c_in + kOsKc - 1) / kOsKc;
  D3B_CUDA(launch_maybe_pdl(spconv_os16_kernel<COUT>, dim3(grid), dim3(Cfg::kThreads), Cfg::kSmemBytes, stream,
                            (const __half*)p->in_hi, (const __half*)p->in_lo, (const int*)nbr, (const unsigned int*)tile_mask,
                            (const int*)n_out, (int)out_cap, (int)p->c_in, n_kb, pack, (int)p->k_vol,
                            (const __half*)p->weight_packed, epi_of(p),
                            (__half*)p->out_hi, (__half*)p->out_lo, p->out_f32, (int*)p->overflow));
  D3B_LAUNCH_CHECK();
  return D3B_OK;
}

template <int COUT>
static int launch_first16(const d3b_conv16_params* p, const int32_t* nbr, const int32_t* n_out, int32_t out_cap,
                          cudaStream_t stream) {
  const size_t smem = (size_t)p->k_vol * p->c_in * COUT * sizeof(float);
  static SmemOptIn optin;
  D3B_REQUIRE(smem <= 160 * 1024, "first-layer sparse conv: weights (%zu bytes) do not fit in shared memory", smem);
  D3B_CUDA(ensure_dynamic_smem(spconv_first16_kernel<COUT>, smem, optin));
  const int grid = grid_for((long long)out_cap * (COUT / 16) * 4, 256, 8);
  spconv_first16_kernel<COUT><<<grid, 256, smem, stream>>>(p->in_f32, nbr, n_out, out_cap, p->c_in, p->k_vol, p->weight,
                                                         epi_of(p), (__half*)p->out_hi, (__half*)p->out_lo, p->out_f32,
                                                         p->overflow);
  D3B_LAUNCH_CHECK();
  return D3B_OK;
}

}  // namespace d3b

using namespace d3b;

extern "C" size_t d3b_conv16_packed_weight_halves(int32_t c_in, int32_t c_out, int32_t k_vol) {
  if (c_in < 1 || c_in > 512 || k_vol < 1 || k_vol > 32) return 0;
  if (!(c_out == 16 || c_out == 32 || c_out == 64 || c_out == 128)) return 0;
  const int pack = os16_pack(c_in, k_vol);
  const int n_kb = pack > 1 ? 1 : (c_in + kOsKc - 1) / kOsKc;
  return (size_t)div_up(k_vol, pack) * n_kb * 2 * c_out * kOsKc;
}

extern "C" int d3b_conv16_pack_weight(const float* weight_dev, int32_t c_in, int32_t c_out, int32_t k_vol,
                                      int32_t w_exp, void* packed_dev, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  D3B_REQUIRE(weight_dev && packed_dev, "d3b_conv16_pack_weight: null argument");
  D3B_REQUIRE(w_exp >= -60 && w_exp <= 60, "d3b_conv16_pack_weight: w_exp %d outside [-60, 60]", w_exp);
  const size_t n = d3b_conv16_packed_weight_halves(c_in, c_out, k_vol);
  if (n == 0) {
    set_error("d3b_conv16_pack_weight: unsupported C_in=%d C_out=%d k_vol=%d", c_in, c_out, k_vol);
    return D3B_ERR_UNSUPPORTED;
  }
  const int pack = os16_pack(c_in, k_vol);
  const int n_kb = pack > 1 ? 1 : (c_in + kOsKc - 1) / kOsKc;
  pack_weight16_kernel<<<grid_for((long long)n, 256), 256, 0, stream>>>(weight_dev, c_in, c_out, k_vol, div_up(k_vol, pack),
                                                                        n_kb, pack, ldexpf(1.0f, w_exp), (__half*)packed_dev);
  D3B_LAUNCH_CHECK();
  return D3B_OK;
}

extern "C" int d3b_sparse_conv16(const int32_t* nbr, const uint32_t* tile_mask, const int32_t* n_out, int32_t out_cap,
                                 const d3b_conv16_params* p, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  D3B_REQUIRE(nbr && n_out && p, "d3b_sparse_conv16: null argument");
  D3B_REQUIRE(p->k_vol >= 1 && p->k_vol <= 32 && out_cap >= 0, "d3b_sparse_conv16: bad shape (k_vol %d)", p->k_vol);
  D3B_REQUIRE((p->out_hi != nullptr) == (p->out_lo != nullptr) && (p->out_hi || p->out_f32),
              "d3b_sparse_conv16: give out_hi + out_lo and/or out_f32");
  D3B_REQUIRE((p->scale == nullptr) == (p->shift == nullptr), "d3b_sparse_conv16: scale and shift go together");
  D3B_REQUIRE((p->residual_hi == nullptr) == (p->residual_lo == nullptr), "d3b_sparse_conv16: residual planes go together");
  if (out_cap == 0) return D3B_OK;
  if (p->in_f32) {      // first layer: fp32 rows, few channels
    D3B_REQUIRE(p->weight && p->c_in >= 1 && p->c_in <= 16, "d3b_sparse_conv16: fp32-input layers need weight and C_in <= 16");
    switch (p->c_out) {
      case 16: return launch_first16<16>(p, nbr, n_out, out_cap, stream);
      case 32: return launch_first16<32>(p, nbr, n_out, out_cap, stream);
      case 64: return launch_first16<64>(p, nbr, n_out, out_cap, stream);
      default:
        set_error("d3b_sparse_conv16: fp32-input layer with C_out=%d (16/32/64 built)", p->c_out);
        return D3B_ERR_UNSUPPORTED;
    }
  }
  D3B_REQUIRE(tile_mask && p->in_hi && p->in_lo && p->weight_packed, "d3b_sparse_conv16: null planes / tile_mask / packed weights");
  if (!os16_shape_ok(p->c_in, p->c_out)) {
    set_error("d3b_sparse_conv16: unsupported C_in=%d C_out=%d", p->c_in, p->c_out);
    return D3B_ERR_UNSUPPORTED;
  }
  switch (p->c_out) {
    case 16: return launch_os16<16>(p, nbr, tile_mask, n_out, out_cap, stream);
    case 32: return launch_os16<32>(p, nbr, tile_mask, n_out, out_cap, stream);
    case 64: return launch_os16<64>(p, nbr, tile_mask, n_out, out_cap, stream);
    default: return launch_os16<128>(p, nbr, tile_mask, n_out, out_cap, stream);
  }
}

extern "C" int d3b_split16(const float* x, int64_t n, void* hi, void* lo, int32_t* overflow, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  D3B_REQUIRE(n >= 0 && (n == 0 || (x && hi && lo)), "d3b_split16: null argument");
  if (n == 0) return D3B_OK;
  split16_kernel<<<grid_for(n, 256), 256, 0, stream>>>(x, n, (__half*)hi, (__half*)lo, overflow);
  D3B_LAUNCH_CHECK();
  return D3B_OK;
}

extern "C" int d3b_merge16(const void* hi, const void* lo, int64_t n, float* x, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  D3B_REQUIRE(n >= 0 && (n == 0 || (x && hi && lo)), "d3b_merge16: null argument");
  if (n == 0) return D3B_OK;
  merge16_kernel<<<grid_for(n, 256), 256, 0, stream>>>((const __half*)hi, (const __half*)lo, n, x);
  D3B_LAUNCH_CHECK();
  return D3B_OK;
}

extern "C" int d3b_sparse_to_bev16(const void* in_hi, const void* in_lo, const float* in_f32, const int32_t* coors,
                                   const int32_t* n_rows, int32_t row_cap, int32_t channels, const int32_t spatial[3],
                                   int32_t batch, void* out_hi, void* out_lo, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  D3B_REQUIRE(coors && n_rows && spatial && out_hi && out_lo, "d3b_sparse_to_bev16: null argument");
  D3B_REQUIRE((in_f32 != nullptr) != (in_hi != nullptr && in_lo != nullptr), "d3b_sparse_to_bev16: give fp32 rows OR both planes");
  D3B_REQUIRE(channels >= 1 && batch >= 1 && row_cap >= 0, "d3b_sparse_to_bev16: bad shape");
  if (row_cap == 0) return D3B_OK;
  sparse_to_bev16_kernel<<<grid_for((long long)row_cap * channels, 256), 256, 0, stream>>>(
      (const __half*)in_hi, (const __half*)in_lo, in_f32, coors, n_rows, row_cap, channels, spatial[0], spatial[1],
      spatial[2], batch, (__half*)out_hi, (__half*)out_lo);
  D3B_LAUNCH_CHECK();
  return D3B_OK;
}

#ifdef D3B_SOFT_TIMEOUT
extern "C" int d3b_debug_fault_spconv16(unsigned int* host8) {
  cudaError_t e = cudaMemcpyFromSymbol(host8, d3b::g_d3b_fault, 32);
  unsigned int zeros[8] = {0};
  if (e == cudaSuccess) e = cudaMemcpyToSymbol(d3b::g_d3b_fault, zeros, 32);
  return (int)e;
}
extern "C" int d3b_debug_cta_ns_spconv16(unsigned long long* host4096) {
  return (int)cudaMemcpyFromSymbol(host4096, d3b::g_d3b_cta_ns, sizeof(unsigned long long) * 4096);
}
extern "C" int d3b_debug_cta_clk_spconv16(long long* host4096) {
  return (int)cudaMemcpyFromSymbol(host4096, d3b::g_d3b_cta_clk, sizeof(long long) * 4096);
}
extern "C" int d3b_debug_trace_spconv16(long long* host, int clear) {
  cudaError_t e = cudaMemcpyFromSymbol(host, d3b::g_d3b_trace, sizeof(long long) * 16 * 512);
  if (e == cudaSuccess && clear) {
    static long long zeros[16 * 512];
    e = cudaMemcpyToSymbol(d3b::g_d3b_trace, zeros, sizeof(zeros));
  }
  return (int)e;
}
#endif
