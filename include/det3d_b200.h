/*
 * det3d_b200 -- C ABI of the B200-native point-cloud inference hot path.
 *
 * Every entry point is `extern "C"`, takes plain pointers / sizes / a CUDA
 * stream (as void*), returns an int status (0 = D3B_OK), never calls exit(),
 * never allocates device memory and never synchronises the host: the caller
 * owns every buffer (sizes come from the *_workspace_bytes() queries) and the
 * data-dependent row counts stay in device memory (`int*` count arguments).
 *
 * The reference (V2AI/Det3D) has no C/FFI plugin boundary for this path; its
 * operator API is Python.  Each entry point below names the reference
 * interface it stands behind (paths relative to the Det3D repository root).
 * INTEGRATION.md shows the ctypes binding a Det3D maintainer would add.
 */
#ifndef DET3D_B200_H_
#define DET3D_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes ------------------------------------------------------- */
#define D3B_OK 0
#define D3B_ERR_INVALID_ARG 1   /* null pointer, negative size, unsupported shape */
#define D3B_ERR_CUDA 2          /* a CUDA runtime call failed: see d3b_last_error() */
#define D3B_ERR_UNSUPPORTED 3   /* channel count / kernel size not built          */
#define D3B_ERR_WORKSPACE 4     /* workspace too small                            */

/* Human-readable text of the last error raised on the calling thread. */
const char* d3b_last_error(void);
/* ABI version; bumped whenever a signature changes. */
int d3b_abi_version(void);
/* Number of kernels this library has launched since load (process-wide);
 * bench.py reports the delta over the timed region as "gpu_launches". */
unsigned long long d3b_launch_count(void);
/* Programmatic dependent launch between consecutive convolution kernels of a stream (default on): the next kernel's
 * launch latency and prologue overlap the previous kernel's tail; results are unaffected. 0 = plain stream order. */
void d3b_set_pdl(int on);
/* Schedule of d3b_bev_conv16 for 3x3 stride-1 layers whose output blocks are 128 channels wide (the RPN blocks of
 * necks/rpn.py:124-142): 0 = pixel-stationary (two M128 x N128 halves per 16 x 16 pixel tile), 1 = channel-stationary
 * (C_out on the TMEM lanes, one M128 x N256 MMA covers the tile), 2 = automatic (default: channel-stationary when the
 * launch has at least two tiles per SM).  Same products in the same order: the results are bit-identical. */
void d3b_set_bev_variant(int variant);
int d3b_get_bev_variant(void);

/* ========================================================================= *
 * 1. Voxelizer
 *    replaces det3d/ops/point_cloud/point_cloud_ops.py:112-184
 *    (points_to_voxel) + :7-55 (_points_to_voxel_reverse_kernel), the
 *    batch-index prepend of det3d/torchie/parallel/collate.py:130-137 and the
 *    per-voxel mean of det3d/models/readers/voxel_encoder.py:206-211.
 * ========================================================================= */
typedef struct {
  float voxel_size[3];   /* x, y, z                                      */
  float range_min[3];    /* x, y, z lower bound of point_cloud_range     */
  int32_t grid[3];       /* x, y, z cells = round((hi - lo) / voxel_size) */
  int32_t ndim;          /* floats per point (>= 3)                      */
  int32_t max_points;    /* max points kept per voxel                    */
  int32_t max_voxels;    /* max voxels per cloud (the reference `break`) */
} d3b_voxel_cfg;

/* Bytes of scratch needed to voxelize `batch` clouds holding n_points_total
 * points in one call. */
size_t d3b_voxelize_workspace_bytes(const d3b_voxel_cfg* cfg, int32_t n_points_total,
                                    int32_t batch);

/* points        [n_total, ndim] f32 device, clouds concatenated
 * cloud_offsets [batch + 1] i32 HOST: cloud b owns points [off[b], off[b+1])
 * voxels        [batch*max_voxels, max_points, ndim] f32 device, or NULL
 * coors         [batch*max_voxels, 4] i32 device (batch, z, y, x)
 * num_points    [batch*max_voxels] i32 device
 * mean_feats    [batch*max_voxels, ndim] f32 device, or NULL
 * voxel_counts  [batch + 1] i32 device: [b] = voxels of cloud b, [batch] = total.
 * Rows of all clouds are written back to back (cloud 0 first); only the
 * first voxel_counts[batch] rows of each output are defined. */
int d3b_voxelize(const d3b_voxel_cfg* cfg, const float* points, const int32_t* cloud_offsets,
                 int32_t batch, float* voxels, int32_t* coors, int32_t* num_points,
                 float* mean_feats, int32_t* voxel_counts, void* workspace,
                 size_t workspace_bytes, void* stream);

/* Multi-sweep ingest (nuScenes): raw sweeps of one sample -> one cloud [n, n_feat + 1] = (x, y, z, .., time lag).
 * replaces read_file / remove_close / read_sweep and the NuScenes branch of LoadPointCloudFromFile.__call__,
 * det3d/datasets/pipelines/loading.py:17-64,98-124.
 * raw            [n_total, raw_stride] f32 device: the sweeps' file contents back to back, key frame first
 * sweep_offsets  [n_sweeps + 1] i32 HOST, in points
 * transforms     [n_sweeps, 16] f64 HOST row-major 4x4 (read where has_transform[s]; may be NULL otherwise)
 * has_transform, filter_close [n_sweeps] u8 HOST, time_lag [n_sweeps] f32 HOST
 * filter_close[s]: drop points with |x| < radius and |y| < radius before the transform (remove_close)
 * out            [out_cap, n_feat + 1] f32 device, n_out [1] i32 device = min(kept points, out_cap); input order kept */
#define D3B_INGEST_MAX_SWEEPS 16
size_t d3b_ingest_workspace_bytes(int32_t n_points_total);
int d3b_ingest_sweeps(const float* raw, const int32_t* sweep_offsets, int32_t n_sweeps, int32_t raw_stride,
                      int32_t n_feat, const double* transforms, const uint8_t* has_transform, const float* time_lag,
                      const uint8_t* filter_close, float radius, float* out, int32_t out_cap, int32_t* n_out,
                      void* workspace, size_t workspace_bytes, void* stream);

/* ========================================================================= *
 * 2. Rulebook (sparse-convolution index maps)
 *    replaces spconv v1.x `get_indice_pairs` as called by SubMConv3d /
 *    SparseConv3d at det3d/models/backbones/scn.py:106-157,323-355
 *    (spconv is an un-vendored dependency of the reference).
 *
 *    The map is output-stationary: nbr[k * row_cap + o] is the input row
 *    feeding output row o through kernel offset k, or -1.  k enumerates
 *    (kz, ky, kx) row-major.  tile_mask[o / 128] has bit k set when any of
 *    the 128 rows of that tile has a neighbour at offset k.
 * ========================================================================= */
typedef struct {
  int32_t spatial[3];   /* D, H, W of the level                              */
  int32_t batch;
  /* level-0 index: open-addressing hash (key -> row)                         */
  uint64_t* hash_keys;  /* [hash_cap] , NULL when the level uses the bitmap   */
  int32_t* hash_vals;   /* [hash_cap]                                         */
  int32_t hash_cap;     /* power of two                                       */
  /* strided-level index: occupancy bitmap + exclusive popcount prefix        */
  uint32_t* bitmap;     /* [n_words]                                          */
  int32_t* word_prefix; /* [n_words]                                          */
  int64_t n_words;
} d3b_site_index;

size_t d3b_rulebook_workspace_bytes(int64_t n_words);

/* Build the level-0 hash index of `coors` ([n,4] b,z,y,x; n read from *n_rows). */
int d3b_index_build_hash(const int32_t* coors, const int32_t* n_rows, int32_t row_cap,
                         d3b_site_index* index, void* stream);

/* Submanifold rulebook: outputs == inputs (same rows, same order).
 * pair_in / pair_out [k_vol, row_cap] + pair_count [k_vol] (all three or none, may be NULL): the same
 * map additionally compacted per offset into (in_row, out_row) lists for D3B_ALGO_TC_PAIRS. */
int d3b_rulebook_subm(const int32_t* coors, const int32_t* n_rows, int32_t row_cap,
                      const d3b_site_index* index, const int32_t ksize[3],
                      int32_t* nbr, uint32_t* tile_mask, int32_t* pair_in, int32_t* pair_out,
                      int32_t* pair_count, void* stream);

/* Strided sparse conv rulebook.  Output sites = every site reachable from an
 * active input, in ascending linear index ((b*D+z)*H+y)*W+x.  Fills
 * out_index (bitmap form), out_coors [out_cap,4], *n_out, nbr, tile_mask. */
int d3b_rulebook_conv(const int32_t* in_coors, const int32_t* n_in, int32_t in_cap,
                      const d3b_site_index* in_index, const int32_t ksize[3],
                      const int32_t stride[3], const int32_t padding[3],
                      d3b_site_index* out_index, int32_t* out_coors, int32_t* n_out,
                      int32_t out_cap, int32_t* nbr, uint32_t* tile_mask, int32_t* pair_in,
                      int32_t* pair_out, int32_t* pair_count, void* workspace,
                      size_t workspace_bytes, void* stream);

/* ========================================================================= *
 * 3. Sparse convolution (gather -> contraction -> fused epilogue)
 *    replaces spconv v1.x `indice_conv` (gather, torch.mm, scatter-add) plus
 *    the BatchNorm1d(eval) / ReLU / residual that follow it in
 *    det3d/models/backbones/scn.py:73-89,106-157.
 *
 *    out[o,:] = act( (sum_k in[nbr[k][o],:] . W[k] + bias) * scale + shift
 *                    + residual[o,:] )
 * ========================================================================= */
#define D3B_ALGO_SIMT 0     /* fp32 FFMA, output-stationary                              */
#define D3B_ALGO_TC 1       /* tcgen05 3xTF32 (fp32-equivalent), output-stationary tiles  */
#define D3B_ALGO_TC_PAIRS 2 /* tcgen05 3xTF32 over compacted rulebook pairs (per offset:  */
                            /* dense chunks of 128 valid pairs), fp32 atomics into feat_out */

typedef struct {
  int32_t c_in, c_out, k_vol;    /* k_vol = kd*kh*kw                          */
  const float* weight;           /* [k_vol, c_in, c_out] f32 (spconv layout)  */
  const float* weight_packed;    /* D3B_ALGO_TC operand image, see below      */
  const float* bias;             /* [c_out] or NULL                           */
  const float* scale;            /* [c_out] or NULL (folded BN)               */
  const float* shift;            /* [c_out] or NULL                           */
  const float* residual;         /* [n_out, c_out] or NULL                    */
  int32_t relu;
  int32_t algo;
  /* D3B_ALGO_TC_PAIRS only -------------------------------------------------- */
  const int32_t* pair_in;        /* [k_vol, out_cap] input row of pair j of offset k (d3b_rulebook_pairs) */
  const int32_t* pair_out;       /* [k_vol, out_cap] output row                                           */
  const int32_t* pair_count;     /* [k_vol] pairs per offset (device)                                      */
  /* epilogue of the PRODUCER layer, applied to the gathered input rows in registers:
   *   x = relu?((x + in_bias) * in_scale + in_shift); NULL pointers = identity.
   * The kernel itself adds raw sums into feat_out (which it zeroes first) and applies
   * nothing else: bias/scale/shift/residual/relu of THIS layer are the consumer's job
   * (next conv's in_* fields, or d3b_feature_epilogue). */
  const float* in_bias;
  const float* in_scale;
  const float* in_shift;
  int32_t in_relu;
  int32_t out_zeroed;            /* nonzero: rows [0, *n_out) of feat_out are already zero (d3b_zero_rows), the
                                    kernel skips its own clearing launch */
} d3b_conv_params;

/* Compact the output-stationary map into per-offset pair lists (spconv's classic rulebook form):
 * for every k, the (in_row, out_row) of each valid nbr[k][o], densely packed; pair_count[k] pairs. */
int d3b_rulebook_pairs(const int32_t* nbr, const int32_t* n_out, int32_t out_cap, int32_t k_vol,
                       int32_t* pair_in, int32_t* pair_out, int32_t* pair_count, void* stream);

/* Clear rows [0, *n_rows) of up to 16 feature buffers that share a row count, in one launch (the accumulation
 * targets of the D3B_ALGO_TC_PAIRS layers of one resolution).  bufs / channels are HOST arrays of `count` entries;
 * channels[i] % 4 == 0. */
int d3b_zero_rows(float* const* bufs, const int32_t* channels, int32_t count, const int32_t* n_rows,
                  int32_t row_cap, void* stream);

/* In-place epilogue over rows [0, *n_rows): x = relu?((x + bias) * scale + shift + residual). */
int d3b_feature_epilogue(float* feat, const int32_t* n_rows, int32_t row_cap, int32_t channels,
                         const float* bias, const float* scale, const float* shift,
                         const float* residual, int32_t relu, void* stream);

/* Size in floats / fill of the tensor-core weight image (hi/lo TF32 split,
 * K-major 128B-swizzled tiles).  Done once per layer at model load. */
size_t d3b_conv_packed_weight_floats(int32_t c_in, int32_t c_out, int32_t k_vol);
int d3b_conv_pack_weight(const float* weight_dev, int32_t c_in, int32_t c_out, int32_t k_vol,
                         float* packed_dev, void* stream);

int d3b_sparse_conv(const float* feat_in, const int32_t* nbr, const uint32_t* tile_mask,
                    const int32_t* n_out, int32_t out_cap, const d3b_conv_params* p,
                    float* feat_out, void* stream);

/* PillarFeatureNet, eval mode, one PFN layer (every Det3D PointPillars config): decoration
 * (x,y,z,.. | xyz - pillar mean | xy - pillar centre), Linear(ndim+5 -> units, no bias), BatchNorm1d folded
 * to scale/shift, ReLU, max over the max_points slots -- padded slots count as all-zero features, as the
 * reference's mask makes them.  replaces det3d/models/readers/pillar_encoder.py:115-155 (+ PFNLayer :33-47).
 * voxels [row_cap, max_points, ndim] f32, num_points [row_cap] i32, coors [row_cap, 4] i32 (b,z,y,x),
 * n_rows device i32 (live rows; the rest of `out` is zero-filled), weight [units, ndim+5] (nn.Linear
 * layout), out [row_cap, units].  x_offset = vx/2 + range_min_x (likewise y).
 * D3B_ERR_UNSUPPORTED unless 3 <= ndim <= 11 and units in {32, 64, 96, 128}. */
int d3b_pillar_features(const float* voxels, const int32_t* num_points, const int32_t* coors,
                        const int32_t* n_rows, int32_t row_cap, int32_t max_points, int32_t ndim,
                        int32_t units, const float* weight, const float* scale, const float* shift,
                        float vx, float vy, float x_offset, float y_offset, float* out, void* stream);

/* The same reader fused with the voxelizer: the pillar's points are fetched through the voxelizer's
 * per-voxel point-index lists, so the [rows, max_points, ndim] voxel tensor is never materialised (call d3b_voxelize
 * with voxels = NULL).  `lists` = d3b_voxelize_point_lists(...) of the workspace that d3b_voxelize just filled:
 * [batch][max_voxels][max_points] int32 indices into `points`, >= 0x7f000000 = empty slot; valid until that workspace is
 * used again.  voxel_counts = d3b_voxelize's per-cloud counts.  Everything else as d3b_pillar_features. */
const int32_t* d3b_voxelize_point_lists(const d3b_voxel_cfg* cfg, int32_t n_points_total, int32_t batch, void* workspace);
int d3b_pillar_features_lists(const float* points, const int32_t* lists, const int32_t* voxel_counts, int32_t batch,
                              int32_t max_voxels, const int32_t* num_points, const int32_t* coors, const int32_t* n_rows,
                              int32_t row_cap, int32_t max_points, int32_t ndim, int32_t units, const float* weight,
                              const float* scale, const float* shift, float vx, float vy, float x_offset, float y_offset,
                              float* out, void* stream);

/* .dense(): rows -> zero-initialised [B, C, D, H, W] (caller zero-fills `out`).
 * replaces SparseConvTensor.dense() at scn.py:192,365. */
int d3b_sparse_to_dense(const float* feat, const int32_t* coors, const int32_t* n_rows,
                        int32_t row_cap, int32_t channels, const int32_t spatial[3],
                        int32_t batch, float* out, void* stream);

/* rows -> zero-initialised channels-last BEV rows [B*H*W, C*D], channel = c*D + z: the same
 * values as .dense().view(B, C*D, H, W) (scn.py:192-195), laid out for the NHWC dense path. */
int d3b_sparse_to_bev_rows(const float* feat, const int32_t* coors, const int32_t* n_rows,
                           int32_t row_cap, int32_t channels, const int32_t spatial[3],
                           int32_t batch, float* out_rows, void* stream);

/* Static rulebook of a dense stride-1 2-D convolution over a [B, H, W] grid (row = (b*H+y)*W+x):
 * nbr[k*n + row] = row of (y + ky - pad_y, x + kx - pad_x) or -1; k = ky*kw + kx.  Lets the
 * dense RPN (det3d/models/necks/rpn.py:124-159) and head 1x1 convs (mg_head.py:198-230) run
 * through d3b_sparse_conv in channels-last layout.  n = B*H*W is also written to n_rows[0..1]. */
int d3b_rulebook_dense2d(int32_t batch, int32_t height, int32_t width, const int32_t ksize[2],
                         const int32_t padding[2], int32_t* nbr, uint32_t* tile_mask,
                         int32_t* n_rows, void* stream);

/* ========================================================================= *
 * 3b. Split-f16 ("FP16x3") convolutions: fp32-equivalent accuracy on the f16 tensor pipe, deterministic.
 *
 *     Activations are carried as TWO f16 planes of the same shape, hi = f16(x) and lo = f16(x - hi) (x = hi + lo
 *     to 22 significant bits, |x| < 65504); weights are split the same way after an exact power-of-two scaling
 *     2^w_exp (undone by `acc_scale` = 2^-w_exp in the epilogue).  The kernels compute hi.hi + hi.lo + lo.hi with
 *     fp32 accumulation.  A result outside the f16 range cannot be carried: the kernels then OR 1 into `*overflow`
 *     (device int, may be NULL) and the caller must fall back to the tf32 path -- nothing saturates silently.
 *     Same reference call sites as section 3 (scn.py:106-157,323-355; necks/rpn.py:82-159; mg_head.py:198-230).
 * ========================================================================= */
typedef struct {
  int32_t c_in, c_out, k_vol;
  const void* in_hi;             /* f16 [rows, c_in] (c_in % 8 == 0)                                          */
  const void* in_lo;
  const float* in_f32;           /* first layer only: fp32 rows [rows, c_in <= 16] (then in_hi = in_lo = NULL)  */
  const float* weight;           /* fp32 [k_vol, c_in, c_out]: used by the fp32-input first layer only          */
  const void* weight_packed;     /* d3b_conv16_pack_weight image                                                */
  float acc_scale;               /* 2^-w_exp (1 for the fp32-input layer)                                       */
  const float* bias;             /* [c_out] or NULL                                                             */
  const float* scale;            /* folded BatchNorm [c_out] or NULL (with shift)                               */
  const float* shift;
  const void* residual_hi;       /* f16 [rows, c_out] planes added before the ReLU, or NULL                     */
  const void* residual_lo;
  int32_t relu;
  void* out_hi;                  /* f16 [rows, c_out] planes (both or neither)                                  */
  void* out_lo;
  float* out_f32;                /* optional fp32 copy of the result [rows, c_out]                              */
  int32_t* overflow;             /* device flag, may be NULL                                                    */
} d3b_conv16_params;

/* Size in halves / fill of the f16 weight image: packed[k][kb][hi|lo][n][64 channels, 128B-swizzled].
 * 0 if the shape is not built (c_out in {16,32,64,128}, c_in <= 512, k_vol <= 32). */
size_t d3b_conv16_packed_weight_halves(int32_t c_in, int32_t c_out, int32_t k_vol);
int d3b_conv16_pack_weight(const float* weight_dev, int32_t c_in, int32_t c_out, int32_t k_vol, int32_t w_exp,
                           void* packed_dev, void* stream);

/* Output-stationary sparse convolution over a rulebook (nbr / tile_mask as in d3b_sparse_conv): no atomics, fixed
 * summation order -> bit-identical results run to run. */
int d3b_sparse_conv16(const int32_t* nbr, const uint32_t* tile_mask, const int32_t* n_out, int32_t out_cap,
                      const d3b_conv16_params* p, void* stream);

/* fp32 <-> plane conversions (API boundaries) and the sparse -> dense NHWC scatter on planes
 * (channel = c*D + z, as d3b_sparse_to_bev_rows; rows given as planes OR as fp32). */
int d3b_split16(const float* x, int64_t n, void* hi, void* lo, int32_t* overflow, void* stream);
int d3b_merge16(const void* hi, const void* lo, int64_t n, float* x, void* stream);
int d3b_sparse_to_bev16(const void* in_hi, const void* in_lo, const float* in_f32, const int32_t* coors,
                        const int32_t* n_rows, int32_t row_cap, int32_t channels, const int32_t spatial[3],
                        int32_t batch, void* out_hi, void* out_lo, void* stream);

/* Dense NHWC convolution on planes [batch, h_in, w_in, c_in] through TMA tensor maps: 3x3 (stride 1 or 2) or 1x1,
 * zero padding `pad`, `groups` weight blocks of c_out (32/64/128) channels each in one launch:
 *   group g -> channel block cg = g % cgroups, sub-pixel ug = g / cgroups, (uy, ux) = (ug / up, ug % up);
 *   conv output pixel (y, x) of group g is written to pixel (y*up + uy, x*up + ux) of the output tensor
 *   [batch, h_out*up, w_out*up, out_channels], channels [out_c0 + cg*c_out, out_c0 + (cg+1)*c_out).
 * cgroups > 1 tiles a wide C_out; up > 1 (with ksize 1) is ConvTranspose2d(kernel = stride = up)
 * (necks/rpn.py:108-122).  weight_packed = the groups' d3b_conv16_pack_weight images back to back;
 * bias / scale / shift hold groups * c_out entries, group-major. */
typedef struct {
  int32_t batch, h_in, w_in, c_in;
  int32_t c_out;                 /* per group */
  int32_t ksize, stride, pad;
  int32_t groups, cgroups, up;
  const void* in_hi;
  const void* in_lo;
  const void* weight_packed;
  float acc_scale;
  const float* bias;
  const float* scale;
  const float* shift;
  int32_t relu;
  int32_t out_channels, out_c0;
  void* out_hi;
  void* out_lo;
  float* out_f32;
  int32_t* overflow;
} d3b_bev16_params;

int d3b_bev_conv16(const d3b_bev16_params* p, void* stream);

/* ========================================================================= *
 * 4. Rotated-box BEV IoU / NMS
 * ========================================================================= */
#define D3B_BOX_XYXYR 0   /* [x1,y1,x2,y2,ry]: det3d/ops/iou3d/src/iou3d_kernel.cu:108-221 */
#define D3B_BOX_XYWLR 1   /* [cx,cy,w,l,r]  : det3d/ops/nms/nms_cpu.py:34-45 + nms_cpu.h:73-169 */
#define D3B_BOX_XYWLR_RRPN 2 /* [cx,cy,w,l,r]: numba RRPN routine, det3d/ops/nms/nms_gpu.py:180-470 */

/* Pairwise rotated IoU (mode 0) or overlap area (mode 1), out [na, nb].
 * replaces boxes_iou_bev_gpu / boxes_overlap_bev_gpu, det3d/ops/iou3d/src/iou3d.cpp:31-71 */
int d3b_boxes_iou_bev(const float* boxes_a, int32_t na, const float* boxes_b, int32_t nb,
                      int32_t mode, float* out, void* stream);

/* RRPN rotated IoU matrix, out[n, k] = f(query[k], boxes[n]); criterion -1: IoU, 0: inter / area(query),
 * 1: inter / area(box), other: intersection area.  replaces rotate_iou_gpu / rotate_iou_gpu_eval,
 * det3d/ops/nms/nms_gpu.py:499-669 (numba.cuda). */
int d3b_rotate_iou_rrpn(const float* boxes, int32_t n, const float* query_boxes, int32_t k,
                        int32_t criterion, float* out, void* stream);

size_t d3b_nms_workspace_bytes(int32_t n_cap);

/* Greedy NMS over boxes already sorted by descending score.
 * fmt = D3B_BOX_XYXYR: suppress when iou >  thresh   (iou3d.cpp:73-120, nms_gpu)
 * fmt = D3B_BOX_XYWLR: suppress when iou >= thresh and the axis-aligned hulls
 *                      overlap                         (nms_cpu.h:73-169)
 * fmt = D3B_BOX_XYWLR_RRPN: suppress when iou > thresh (rotate_nms_gpu, nms_gpu.py:411-496)
 * n_boxes may be a device count (n_boxes_dev != NULL, bounded by n_cap).
 * keep_idx [min(n_cap, max_keep)] i64 device receives kept positions in
 * ascending order, keep_count [1] i32 device their number (<= max_keep). */
int d3b_rotate_nms(const float* boxes, int32_t n_cap, const int32_t* n_boxes_dev, int32_t fmt,
                   float thresh, int32_t max_keep, int64_t* keep_idx, int32_t* keep_count,
                   void* workspace, size_t workspace_bytes, void* stream);

/* Axis-aligned variants, boxes [n,5] = (x1, y1, x2, y2, ignored), suppress when iou > thresh.
 * mode = D3B_AA_IOU3D: plain extents       -- nms_normal_gpu, iou3d.cpp:123-170 / iou3d_kernel.cu:295-303
 * mode = D3B_AA_PIXEL: "+1" pixel extents  -- numba nms_gpu behind box_torch_ops.nms
 *                      (det3d/ops/nms/nms_gpu.py:22-33,129-166, core/bbox/box_torch_ops.py:506-525) */
#define D3B_AA_IOU3D 0
#define D3B_AA_PIXEL 1
int d3b_normal_nms(const float* boxes, int32_t n_cap, const int32_t* n_boxes_dev, int32_t mode, float thresh,
                   int32_t max_keep, int64_t* keep_idx, int32_t* keep_count, void* workspace,
                   size_t workspace_bytes, void* stream);

/* ========================================================================= *
 * 5. Detection post-processing of one task head (whole batch, fixed shapes)
 *    replaces MultiGroupHead.get_task_detections, det3d/models/bbox_heads/mg_head.py:805-1085
 *    (use_multi_class_nms = False), incl. second_box_decode (core/bbox/box_torch_ops.py:80-148)
 *    and rotate_nms (:528-549).  Head tensors are addressed as rows: element (b, cell, j) lives
 *    at ptr[(b*hw + cell) * row_stride + col0 + j], so both NHWC-permuted tensors and column
 *    slices of a fused [B*H*W, 32] row buffer work without a copy.
 * ========================================================================= */
typedef struct {
  const float* cls;  int32_t cls_row_stride, cls_col0;   /* logits, na*n_cls per cell          */
  const float* box;  int32_t box_row_stride, box_col0;   /* encodings, na*code per cell        */
  const float* dir;  int32_t dir_row_stride, dir_col0;   /* direction logits, na*2, or NULL     */
  const float* anchors;                                   /* [hw*na, nd], shared by the batch    */
  int32_t batch, hw, na, n_cls, code, nd;
  int32_t vec_encode, smooth_dim, norm_velo;              /* box coder flags                     */
  int32_t use_rotate_nms, pre_max, post_max;              /* test_cfg.nms                        */
  float nms_iou_threshold, score_threshold, direction_offset;
  float post_center_range[6]; int32_t has_range;
  int32_t label_offset;                                   /* added to the class index            */
} d3b_predict_params;

size_t d3b_predict_workspace_bytes(const d3b_predict_params* p);

/* packed [batch, packed_rows_per_sample, nd+3] f32: rows [row_offset, row_offset+post_max) of every
 * sample receive box[nd], score, label, valid(0/1); rows beyond the kept count are zero.
 * keep_counts [batch] i32 (optional) = boxes surviving NMS (before the range mask). */
int d3b_predict_task(const d3b_predict_params* p, float* packed, int32_t packed_rows_per_sample,
                     int32_t row_offset, int32_t* keep_counts, void* workspace,
                     size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DET3D_B200_H_ */
