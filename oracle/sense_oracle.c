/* sense_oracle.c - CPU restatement of the local-map ACQUISITION stage of mapping_util (SURVEY.md 8(f) row 4,
 * the part in front of map_oracle.c): crop of the environment grid around the agent, 360 degree / limited
 * field-of-view ray-cast clearing, merge with the grid kept from the previous update.
 *
 * TEST INFRASTRUCTURE ONLY: imported by tests/ and bench legs; the product never links or calls it.
 *
 * Restates, in plain C and in the reference's own (sequential) order:
 *   sense_frame      origin / dimension / start index      mapping_util/src/map_builder.cpp:89-118
 *   sense_crop       sub-grid of the environment message   map_builder.cpp:120-153
 *   sense_clear_line MapBuilder::ClearLine                 map_builder.cpp:367-432
 *   sense_raycast    MapBuilder::RaycastAndClear           map_builder.cpp:280-329
 *   sense_merge      MapBuilder::MergeVoxelGrids           map_builder.cpp:242-278
 *   sense_center     MapBuilder::ClearVoxelsCenter         map_builder.cpp:434-447
 *   sense_update     the sequence of map_builder.cpp:80-205 for one agent
 *
 * Parity: PINNED.  The reference's map-builder node itself - mapping_util/src/map_builder.cpp with
 * path_finding_util/src/path_tools.cpp, voxel_grid_util/src/raycast.cpp and voxel_grid.cpp, compiled UNMODIFIED from
 * /root/reference against stand-in headers for rclcpp / tf2 / the message classes / Eigen (oracle/ref_shim/, none of it
 * reference code) into oracle/_ref/libref_map.so - is driven through MapBuilder::EnvironmentVoxelGridCallback by
 * oracle/ref_wrap_map.cpp.  sense_update reproduces its voxel_grid_curr_ and origin byte for byte, first and follow-up
 * updates, 360 degree / limited field of view / known map: live where the reference is present and through the committed
 * fixture tests/golden/mapbuilder_ref.npz (tests/golden/make_mapbuilder_golden.py) elsewhere (tests/test_sense_oracle.py).
 * One caveat: the field-of-view test's dot products run through the Eigen stand-in, whose evaluation order
 * ((a0 b0 + a1 b1) + a2 b2) is assumed to be real Eigen's; everything else is integer / IEEE-exact logic.
 * Grids are [dz][dy][dx] int8, x fastest: 0 free, 100 occupied, -1 unknown.
 */
#include "reftraj_oracle.c" /* rt_raycast, vg_inside, vg_get */

typedef struct sense_params {
  double voxel;     /* voxel_grid.voxel_size of the environment message */
  double range[3];  /* voxel_grid_range_ */
  int32_t free_grid;   /* free_grid_: 1 = no ray casting, unknown -> free */
  int32_t limited_fov; /* limited_fov_ */
  double cos_half_fov_x, cos_half_fov_y; /* cos(fov_x_ / 2), cos(fov_y_ / 2) */
} sense_params;

/* map_builder.cpp:89-118.  origin_env: origin of the environment grid. */
void sense_frame(const sense_params* P, const double origin_env[3], const double pos[3], double origin[3], int32_t dim[3],
                 int32_t start_idx[3]) {
  for (int a = 0; a < 3; ++a) {
    double o = pos[a] - P->range[a] / 2;
    o = round((o - origin_env[a]) / P->voxel) * P->voxel + origin_env[a];
    origin[a] = o;
    dim[a] = (int32_t)floor(P->range[a] / P->voxel);
    start_idx[a] = (int32_t)round((o - origin_env[a]) / P->voxel);
  }
}

/* map_builder.cpp:120-153 */
void sense_crop(const sense_params* P, const int8_t* env, const int32_t dim_env[3], const int32_t dim[3], const int32_t start_idx[3],
                int8_t* vg) {
  for (int i = 0; i < dim[0]; ++i)
    for (int j = 0; j < dim[1]; ++j)
      for (int k = 0; k < dim[2]; ++k) {
        const int ie = i + start_idx[0], je = j + start_idx[1], ke = k + start_idx[2];
        int8_t v = vg_inside(dim_env, ie, je, ke) ? env[ie + dim_env[0] * je + dim_env[0] * dim_env[1] * ke] : -1;
        if (P->free_grid) {
          if (v == -1) v = 0;
        } else if (v == 0)
          v = -1;
        vg[i + dim[0] * j + dim[0] * dim[1] * k] = v;
      }
}

static void set_voxel(int8_t* data, const int32_t dim[3], int x, int y, int z, int8_t v) { /* SetVoxelInt, voxel_grid.cpp:84-93 */
  if (vg_inside(dim, x, y, z)) data[x + y * dim[0] + z * dim[0] * dim[1]] = v;
}

static double dot3(const double* a, const double* b) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }
static void normalize3(double* v) { /* Eigen normalize(): v /= sqrt(squaredNorm) when that is > 0 */
  const double z = dot3(v, v);
  if (z > 0) {
    const double n = sqrt(z);
    v[0] /= n, v[1] /= n, v[2] /= n;
  }
}

/* map_builder.cpp:367-432.  rot: rot_mat_cam_ row-major 3x3 (columns are the body axes); may be NULL without limited_fov. */
int sense_clear_line(const sense_params* P, const int8_t* vg, int8_t* vg_final, const int32_t dim[3], const double start[3],
                     const double end[3], const double* rot) {
  if (P->limited_fov) {
    const double xb[3] = {rot[0], rot[3], rot[6]}, yb[3] = {rot[1], rot[4], rot[7]}, zb[3] = {rot[2], rot[5], rot[8]};
    const double dir[3] = {end[0] - start[0], end[1] - start[1], end[2] - start[2]};
    const double dy = dot3(dir, yb), dz = dot3(dir, zb);
    double dir_xz[3], dir_xy[3];
    for (int a = 0; a < 3; ++a) dir_xz[a] = dir[a] - dy * yb[a], dir_xy[a] = dir[a] - dz * yb[a]; /* y_b twice, as written (:385) */
    normalize3(dir_xz);
    normalize3(dir_xz); /* the second normalize() is on dir_xz as well (:386); dir_xy stays un-normalised */
    if (!(dot3(dir_xz, xb) > P->cos_half_fov_y && dot3(dir_xy, xb) > P->cos_half_fov_x)) return 0;
  }
  static __thread double visited[3 * 1600];
  double col[3];
  const double ddx = start[0] - end[0], ddy = start[1] - end[1], ddz = start[2] - end[2];
  const double max_dist = sqrt((ddx * ddx + ddy * ddy) + ddz * ddz);
  int n = rt_raycast(vg, dim, start, end, max_dist, visited, 1590, col);
  if (n < 0) return -1; /* the reference throws (more than 1500 voxels on one ray) */
  if (col[0] == -1) { /* line clear: the end point closes the list (:404-405) */
    visited[3 * n] = end[0], visited[3 * n + 1] = end[1], visited[3 * n + 2] = end[2];
    ++n;
  } else {
    const double lp[3] = {(end[0] - start[0]) * 1e-7 + col[0], (end[1] - start[1]) * 1e-7 + col[1], (end[2] - start[2]) * 1e-7 + col[2]};
    set_voxel(vg_final, dim, (int)lp[0], (int)lp[1], (int)lp[2], 100);
  }
  for (int i = 0; i < n - 1; ++i)
    set_voxel(vg_final, dim, (int)((visited[3 * i] + visited[3 * i + 3]) / 2.0), (int)((visited[3 * i + 1] + visited[3 * i + 4]) / 2.0),
              (int)((visited[3 * i + 2] + visited[3 * i + 5]) / 2.0), 0);
  return 0;
}

/* map_builder.cpp:280-329: rays to every border voxel, floor and ceiling first, then the y walls, then the x walls */
int sense_raycast(const sense_params* P, const int8_t* vg, int8_t* vg_final, const int32_t dim[3], const double start[3], const double* rot) {
  memset(vg_final, -1, (size_t)dim[0] * dim[1] * dim[2]);
  const int kv[2] = {0, dim[2] - 1}, jv[2] = {0, dim[1] - 1}, iv[2] = {0, dim[0] - 1};
  int rc = 0;
  for (int i = 0; i < dim[0]; ++i)
    for (int j = 0; j < dim[1]; ++j)
      for (int s = 0; s < 2; ++s) {
        const double end[3] = {i + 0.5, j + 0.5, kv[s] + 0.5};
        rc |= sense_clear_line(P, vg, vg_final, dim, start, end, rot);
      }
  for (int i = 0; i < dim[0]; ++i)
    for (int k = 0; k < dim[2]; ++k)
      for (int s = 0; s < 2; ++s) {
        const double end[3] = {i + 0.5, jv[s] + 0.5, k + 0.5};
        rc |= sense_clear_line(P, vg, vg_final, dim, start, end, rot);
      }
  for (int j = 0; j < dim[1]; ++j)
    for (int k = 0; k < dim[2]; ++k)
      for (int s = 0; s < 2; ++s) {
        const double end[3] = {iv[s] + 0.5, j + 0.5, k + 0.5};
        rc |= sense_clear_line(P, vg, vg_final, dim, start, end, rot);
      }
  return rc;
}

/* map_builder.cpp:242-278: unknown voxels of the new grid take the old grid's value at the same world position */
void sense_merge(const int8_t* vg_old, const int32_t dim_old[3], const double origin_old[3], int8_t* vg_new, const int32_t dim[3],
                 const double origin_new[3], double voxel) {
  int off[3];
  for (int a = 0; a < 3; ++a) off[a] = (int)round((origin_new[a] - origin_old[a]) / voxel);
  for (int i = 0; i < dim[0]; ++i)
    for (int j = 0; j < dim[1]; ++j)
      for (int k = 0; k < dim[2]; ++k) {
        int8_t* v = &vg_new[i + dim[0] * j + dim[0] * dim[1] * k];
        if (*v == -1) *v = (int8_t)vg_get(vg_old, dim_old, i + off[0], j + off[1], k + off[2]);
      }
}

/* map_builder.cpp:434-447 on an all-unknown grid (:160-167) */
void sense_center(int8_t* vg, const int32_t dim[3], const double pos_local[3]) {
  memset(vg, -1, (size_t)dim[0] * dim[1] * dim[2]);
  const int im = (int)floor(pos_local[0]), jm = (int)floor(pos_local[1]), km = (int)floor(pos_local[2]);
  for (int i = im - 2; i <= im + 2; ++i)
    for (int j = jm - 2; j <= jm + 2; ++j)
      for (int k = km - 2; k <= km + 2; ++k) set_voxel(vg, dim, i, j, k, 0);
}

/* One map update of one agent (map_builder.cpp:80-205).  old_grid / old_origin: voxel_grid_curr_ of the previous
 * update (same dimension), have_old = 0 on the first update.  Writes the new voxel_grid_curr_ and its origin. */
int sense_update(const sense_params* P, const int8_t* env, const int32_t dim_env[3], const double origin_env[3], const double pos[3],
                 const double* rot, const int8_t* old_grid, const double* old_origin, int have_old, int8_t* out, double origin_out[3]) {
  int32_t dim[3], start_idx[3];
  sense_frame(P, origin_env, pos, origin_out, dim, start_idx);
  const size_t n = (size_t)dim[0] * dim[1] * dim[2];
  if (P->free_grid) {
    sense_crop(P, env, dim_env, dim, start_idx, out);
    return 0;
  }
  int8_t* vg = (int8_t*)malloc(n);
  int8_t* first = NULL;
  sense_crop(P, env, dim_env, dim, start_idx, vg);
  double pl[3];
  for (int a = 0; a < 3; ++a) pl[a] = (pos[a] - origin_out[a]) / P->voxel; /* GetCoordLocal, voxel_grid.cpp:137-142 */
  if (!have_old) {
    first = (int8_t*)malloc(n);
    sense_center(first, dim, pl);
    old_grid = first, old_origin = origin_out;
  }
  const int rc = sense_raycast(P, vg, out, dim, pl, rot);
  sense_merge(old_grid, dim, old_origin, out, dim, origin_out, P->voxel);
  free(vg);
  free(first);
  return rc;
}

typedef struct sense_job {
  const sense_params* P;
  int n, n_threads, tid;
  const int8_t* env;
  const int32_t* dim_env;
  const double *origin_env, *pos, *rot, *old_origin;
  const int8_t* old_grids;
  const uint8_t* have_old;
  size_t stride;
  int8_t* out;
  double* origin_out;
  int rc;
} sense_job;

static void* sense_worker(void* arg) {
  sense_job* J = (sense_job*)arg;
  for (int a = J->tid; a < J->n; a += J->n_threads)
    J->rc |= sense_update(J->P, J->env, J->dim_env, J->origin_env, J->pos + 3 * a, J->rot ? J->rot + 9 * a : NULL,
                          J->old_grids ? J->old_grids + J->stride * a : NULL, J->old_origin ? J->old_origin + 3 * a : NULL,
                          J->have_old ? J->have_old[a] : 0, J->out + J->stride * a, J->origin_out + 3 * a);
  return NULL;
}

/* n agents in one environment grid; grids at a stride of `stride` voxels */
int sense_update_batch(const sense_params* P, int n, const int8_t* env, const int32_t* dim_env, const double* origin_env, const double* pos,
                       const double* rot, const int8_t* old_grids, const double* old_origin, const uint8_t* have_old, size_t stride,
                       int8_t* out, double* origin_out, int n_threads) {
  if (n_threads < 1) n_threads = 1;
  if (n_threads > 256) n_threads = 256;
  pthread_t th[256];
  sense_job jobs[256];
  for (int t = 0; t < n_threads; ++t) {
    jobs[t] = (sense_job){P, n, n_threads, t, env, dim_env, origin_env, pos, rot, old_origin, old_grids, have_old, stride, out, origin_out, 0};
    pthread_create(&th[t], NULL, sense_worker, &jobs[t]);
  }
  int rc = 0;
  for (int t = 0; t < n_threads; ++t) pthread_join(th[t], NULL), rc |= jobs[t].rc;
  return rc;
}
