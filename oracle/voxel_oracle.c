/*
 * TEST INFRASTRUCTURE -- CPU oracle, not product code.
 *
 * Plain-C restatement of the reference voxelizer
 *   det3d/ops/point_cloud/point_cloud_ops.py:7-55   (_points_to_voxel_reverse_kernel)
 *   det3d/ops/point_cloud/point_cloud_ops.py:112-184 (points_to_voxel)
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this.  Pinned against the reference numba function
 * itself through tests/golden/voxel_*.npz (made by tests/golden/make_golden.py).
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off: fp32 sub, IEEE
 * division and floorf exactly as numba evaluates point_cloud_ops.py:36).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* grid = round((hi - lo) / vs) in fp32, half-to-even like np.round (:26-29). */
void oracle_voxel_grid(const float* vs, const float* range, int32_t* grid) {
  for (int j = 0; j < 3; ++j) {
    float g = (range[3 + j] - range[j]) / vs[j];
    grid[j] = (int32_t)nearbyintf(g);
  }
}

/*
 * dense_map: int32[gz*gy*gx] scratch pre-filled with -1, or NULL to allocate
 * and fill it per call the way point_cloud_ops.py:150 does (that fill is what
 * the reference spends its time on, so the "full call" CPU baseline passes NULL).
 * Outputs must be zero-initialised by the caller (:149-154).
 * Returns voxel_num (:181-184 slice bound), or -1 on allocation failure.
 */
int32_t oracle_points_to_voxel(const float* points, int32_t n, int32_t ndim, const float* vs,
                               const float* range, int32_t max_points, int32_t max_voxels,
                               float* voxels, int32_t* coors, int32_t* num_points_per_voxel,
                               int32_t* dense_map) {
  int32_t grid[3];
  oracle_voxel_grid(vs, range, grid);
  const size_t cells = (size_t)grid[0] * grid[1] * grid[2];
  int32_t* map = dense_map;
  if (map == NULL) {
    map = (int32_t*)malloc(cells * sizeof(int32_t));
    if (!map) return -1;
    for (size_t i = 0; i < cells; ++i) map[i] = -1;              /* -np.ones(shape, int32), :150 */
  }
  int32_t voxel_num = 0;
  for (int32_t i = 0; i < n; ++i) {                                /* :33 */
    int32_t c[3];
    int failed = 0;
    for (int j = 0; j < 3; ++j) {                                  /* :35, ndim fixed to 3 (:24) */
      float f = floorf((points[(size_t)i * ndim + j] - range[j]) / vs[j]); /* :36 */
      if (f < 0 || f >= (float)grid[j]) { failed = 1; break; }     /* :37-39 */
      c[j] = (int32_t)f;
    }
    if (failed) continue;
    /* coor = (z, y, x) (:40); map indexed [z][y][x] */
    const size_t cell = ((size_t)c[2] * grid[1] + c[1]) * grid[0] + c[0];
    int32_t vid = map[cell];
    if (vid == -1) {                                               /* :44 */
      vid = voxel_num;
      if (voxel_num >= max_voxels) break;                          /* :46-47: BREAK, not continue */
      voxel_num += 1;
      map[cell] = vid;
      coors[vid * 3 + 0] = c[2];
      coors[vid * 3 + 1] = c[1];
      coors[vid * 3 + 2] = c[0];
    }
    const int32_t num = num_points_per_voxel[vid];
    if (num < max_points) {                                        /* :51-54 */
      memcpy(voxels + ((size_t)vid * max_points + num) * ndim, points + (size_t)i * ndim,
             sizeof(float) * ndim);
      num_points_per_voxel[vid] = num + 1;
    }
  }
  if (dense_map == NULL) {
    free(map);
  } else {
    /* leave the caller's scratch all -1 again (sparse reset) */
    for (int32_t v = 0; v < voxel_num; ++v)
      map[((size_t)coors[v * 3] * grid[1] + coors[v * 3 + 1]) * grid[0] + coors[v * 3 + 2]] = -1;
  }
  return voxel_num;
}
