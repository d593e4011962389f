// TEST INFRASTRUCTURE - C entry points around the reference's own voxel grid and ray caster.
//
// Linked with /root/reference/voxel_grid_util/src/{voxel_grid,raycast}.cpp (compiled unmodified from where
// they lie, against the <Eigen/Dense> stand-in in oracle/ref_shim/) into oracle/_ref/libref_voxel.so.
// Used by tests/ to pin the ray-casting part of oracle/reftraj_oracle.c (voxel_grid_util::Raycast,
// raycast.cpp:21-186, the core of path_finding_util::IsLineClear, path_tools.cpp:148-180) and the grid
// post-processing of oracle/map_oracle.c (VoxelGrid::CreateMask / InflateObstacles / CreatePotentialField).
#include <stdint.h>
#include <string.h>

#include <vector>

#include "raycast.hpp"

extern "C" {

// Raycast in LOCAL voxel coordinates (what GetCoordLocal returns) through the grid `data`
// ([dim z][dim y][dim x] int8, x fastest).  visited receives up to cap points (the reference's output
// vector, in order); returns their number.  collision receives collision_pt ((-1,-1,-1) when clear).
int ref_raycast(const int8_t* data, const int32_t dim[3], const double start[3], const double end[3], double max_dist,
                double* visited, int cap, double collision[3]) {
  const size_t n = (size_t)dim[0] * dim[1] * dim[2];
  std::vector<voxel_grid_util::voxel_data_type> grid(data, data + n);
  Eigen::Vector3d origin(0, 0, 0);
  Eigen::Vector3i d(dim[0], dim[1], dim[2]);
  voxel_grid_util::VoxelGrid vg(origin, d, 1.0, grid);
  Eigen::Vector3d s(start[0], start[1], start[2]), e(end[0], end[1], end[2]), col(-1, -1, -1);
  const std::vector<Eigen::Vector3d> out = voxel_grid_util::Raycast(s, e, col, vg, max_dist, false);
  for (size_t i = 0; i < out.size() && (int)i < cap; ++i)
    for (int a = 0; a < 3; ++a) visited[3 * i + a] = out[i](a);
  for (int a = 0; a < 3; ++a) collision[a] = col(a);
  return (int)out.size();
}

// VoxelGrid::InflateObstacles followed (when potential_dist > 0) by VoxelGrid::CreatePotentialField
// (voxel_grid.cpp:251-298), the grid post-processing of mapping_util (map_builder.cpp:211-216), in place.
void ref_inflate_and_potential(int8_t* data, const int32_t dim[3], double vox, double inflation_dist, double potential_dist,
                               int potential_pow) {
  const size_t n = (size_t)dim[0] * dim[1] * dim[2];
  std::vector<voxel_grid_util::voxel_data_type> grid(data, data + n);
  Eigen::Vector3d origin(0, 0, 0);
  Eigen::Vector3i d(dim[0], dim[1], dim[2]);
  voxel_grid_util::VoxelGrid vg(origin, d, vox, grid);
  if (inflation_dist > 0) vg.InflateObstacles(inflation_dist);
  if (potential_dist > 0) vg.CreatePotentialField(potential_dist, potential_pow);
  const std::vector<voxel_grid_util::voxel_data_type> out = vg.GetData();
  memcpy(data, out.data(), n);
}

// VoxelGrid::CreateMask (voxel_grid.cpp:192-226): offsets [cap][3] and values; returns the count.
int ref_create_mask(double vox, double mask_dist, double pow, int32_t* offsets, int8_t* values, int cap) {
  Eigen::Vector3d origin(0, 0, 0);
  Eigen::Vector3i d(1, 1, 1);
  voxel_grid_util::VoxelGrid vg(origin, d, vox, true);
  const auto mask = vg.CreateMask(mask_dist, pow);
  for (size_t i = 0; i < mask.size() && (int)i < cap; ++i) {
    for (int a = 0; a < 3; ++a) offsets[3 * i + a] = mask[i].first(a);
    values[i] = mask[i].second;
  }
  return (int)mask.size();
}
}
