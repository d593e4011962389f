/* map_oracle.c - CPU restatement of the local-map post-processing (the part every
 * replanning step runs on the agent's voxel grid before corridor and reference trajectory are computed).
 *
 * TEST INFRASTRUCTURE ONLY: imported by tests/ and bench legs; the product never links or calls it.
 *
 * Restates, in plain C:
 *   map_create_mask        VoxelGrid::CreateMask             voxel_grid_util/src/voxel_grid.cpp:192-226
 *   map_uncertain          MapBuilder::SetUncertainToUnknown mapping_util/src/map_builder.cpp:331-365
 *   map_inflate            VoxelGrid::InflateObstacles       voxel_grid.cpp:251-277
 *   map_potential          VoxelGrid::CreatePotentialField   voxel_grid.cpp:279-298
 *   map_process            the sequence of map_builder.cpp:207-216
 *
 * Parity: map_create_mask, map_inflate and map_potential are PINNED - tests/test_map_oracle.py compares them
 * byte for byte with the reference's own VoxelGrid (voxel_grid.cpp compiled unmodified into
 * oracle/_ref/libref_voxel.so).  map_uncertain lives in the ROS2 node map_builder.cpp; it is pinned through the grid that
 * node PUBLISHES (map_builder.cpp compiled unmodified on stand-in ROS headers into oracle/_ref/libref_map.so): map_process
 * reproduces it byte for byte (tests/test_sense_oracle.py, fixture tests/golden/mapbuilder_ref.npz).  Grids are [dz][dy][dx] int8, x fastest: 0 free, 100 occupied, -1 unknown, 1..99 potential.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MAP_OCC 100
#define MAP_UNK (-1)

/* offsets [cap][3] and values of the stencil; returns the count (may exceed cap: then only cap are written) */
int map_create_mask(double vox, double mask_dist, double power, int32_t* offsets, int8_t* values, int cap) {
  const int rn = (int)ceil(mask_dist / vox);
  int n = 0;
  if (!(mask_dist > 0)) return 0;
  for (int x = -rn; x <= rn; ++x)
    for (int y = -rn; y <= rn; ++y)
      for (int z = -rn; z <= rn; ++z) {
        const double d = hypot(hypot(x, y), z);
        if (fabs(d - 1) * vox >= mask_dist) continue; /* one voxel is taken off the distance (:207-210) */
        const double h = 100.0 * pow((1 - (double)hypot(hypot(x, y), z) / (rn + 1)), power);
        if (h > 1e-3) {
          if (n < cap) offsets[3 * n] = x, offsets[3 * n + 1] = y, offsets[3 * n + 2] = z, values[n] = (int8_t)h;
          ++n;
        }
      }
  return n;
}

static size_t at(const int32_t dim[3], int x, int y, int z) { return (size_t)x + (size_t)y * dim[0] + (size_t)z * dim[0] * dim[1]; }
static int inside(const int32_t dim[3], int x, int y, int z) { return x >= 0 && y >= 0 && z >= 0 && x < dim[0] && y < dim[1] && z < dim[2]; }

/* SetUncertainToUnknown: every unknown voxel of the interior makes the cube around it unknown, occupied voxels
 * excepted; reads `in`, writes `out` (the reference works on a copy too) */
void map_uncertain(const int8_t* in, int8_t* out, const int32_t dim[3], double vox, double inflation_dist) {
  const int c = (int)ceil(inflation_dist / vox);
  memcpy(out, in, (size_t)dim[0] * dim[1] * dim[2]);
  for (int i = c; i < dim[0] - c; ++i)
    for (int j = c; j < dim[1] - c; ++j)
      for (int k = c; k < dim[2] - c; ++k) {
        if (in[at(dim, i, j, k)] != MAP_UNK) continue;
        for (int a = -c; a <= c; ++a)
          for (int b = -c; b <= c; ++b)
            for (int d = -c; d <= c; ++d)
              if (inside(dim, i + a, j + b, k + d) && in[at(dim, i + a, j + b, k + d)] != MAP_OCC)
                out[at(dim, i + a, j + b, k + d)] = MAP_UNK;
      }
}

/* InflateObstacles: the voxels occupied BEFORE the call stamp the mask as occupied */
void map_inflate(int8_t* data, const int32_t dim[3], double vox, double inflation_dist) {
  int32_t off[3 * 4096];
  int8_t val[4096];
  const int nm = map_create_mask(vox, inflation_dist, 1, off, val, 4096);
  const size_t n = (size_t)dim[0] * dim[1] * dim[2];
  int8_t* src = (int8_t*)malloc(n);
  memcpy(src, data, n);
  for (int x = 0; x < dim[0]; ++x)
    for (int y = 0; y < dim[1]; ++y)
      for (int z = 0; z < dim[2]; ++z) {
        if (src[at(dim, x, y, z)] != MAP_OCC) continue;
        for (int m = 0; m < nm && m < 4096; ++m)
          if (inside(dim, x + off[3 * m], y + off[3 * m + 1], z + off[3 * m + 2]))
            data[at(dim, x + off[3 * m], y + off[3 * m + 1], z + off[3 * m + 2])] = MAP_OCC;
      }
  free(src);
}

/* CreatePotentialField: occupied voxels raise their surroundings to the mask value; unknown voxels stay unknown.
 * Done in place like the reference: no mask value other than the centre's reaches 100, so the set of occupied
 * voxels does not change while it is being scanned. */
void map_potential(int8_t* data, const int32_t dim[3], double vox, double potential_dist, int power) {
  int32_t off[3 * 4096];
  int8_t val[4096];
  const int nm = map_create_mask(vox, potential_dist, power, off, val, 4096);
  for (int x = 0; x < dim[0]; ++x)
    for (int y = 0; y < dim[1]; ++y)
      for (int z = 0; z < dim[2]; ++z) {
        if (data[at(dim, x, y, z)] != MAP_OCC) continue;
        for (int m = 0; m < nm && m < 4096; ++m) {
          const int a = x + off[3 * m], b = y + off[3 * m + 1], c = z + off[3 * m + 2];
          if (!inside(dim, a, b, c)) continue;
          int8_t* t = &data[at(dim, a, b, c)];
          if (*t != MAP_UNK && val[m] > *t) *t = val[m];
        }
      }
}

void map_process(const int8_t* in, int8_t* out, const int32_t dim[3], double vox, double inflation_dist, double potential_dist,
                 int power) {
  map_uncertain(in, out, dim, vox, inflation_dist);
  map_inflate(out, dim, vox, inflation_dist);
  map_potential(out, dim, vox, potential_dist, power);
}

typedef struct {
  const int8_t* in;
  int8_t* out;
  const int32_t* dims;
  size_t stride;
  int n, tid, nt, power;
  double vox, infl, pot;
} map_job;
static void* map_worker(void* arg) {
  const map_job* J = (const map_job*)arg;
  for (int i = J->tid; i < J->n; i += J->nt)
    map_process(J->in + (size_t)i * J->stride, J->out + (size_t)i * J->stride, J->dims + 3 * i, J->vox, J->infl, J->pot, J->power);
  return 0;
}
int map_process_batch(int n, const int8_t* in, int8_t* out, size_t stride, const int32_t* dims, double vox, double inflation_dist,
                      double potential_dist, int power, int n_threads) {
  if (n_threads < 1) n_threads = 1;
  if (n_threads > 256) n_threads = 256;
  map_job jobs[256];
  pthread_t th[256];
  for (int t = 0; t < n_threads; ++t) {
    map_job j = {in, out, dims, stride, n, t, n_threads, power, vox, inflation_dist, potential_dist};
    jobs[t] = j;
  }
  for (int t = 1; t < n_threads; ++t) pthread_create(&th[t], 0, map_worker, &jobs[t]);
  map_worker(&jobs[0]);
  for (int t = 1; t < n_threads; ++t) pthread_join(th[t], 0);
  return 0;
}
