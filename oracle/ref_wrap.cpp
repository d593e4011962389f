// TEST INFRASTRUCTURE - C entry points around the reference's own corridor generator.
//
// Linked with /root/reference/convex_decomp_util/src/convex_decomp.cpp (compiled unmodified from
// where it lies, against the Eigen stand-in in oracle/ref_shim/) into oracle/_ref/libref_corridor.so.
// Used by tests/ and tests/golden/make_corridor_golden.py to pin oracle/corridor_oracle.c and the CUDA
// kernel against the real GetPolyOcta3D / GetPolyOcta3DNew (convex_decomp.cpp:5-376, :211-...).
#include <stdint.h>
#include <string.h>

#include <vector>

#include "convex_decomp.hpp"

extern "C" {

// Runs the reference decomposition on a copy of `data` ([dim z][dim y][dim x] int8, x fastest, as
// voxel_grid_util lays it out) and returns the number of hyperplanes; points / normals receive
// (p_, n_) of each in the reference's order (chamfers first, then the six faces).  data_out (may be
// null) receives the grid with the voxels of the convex set marked `conv`.
int ref_get_poly_octa_3d(const int32_t seed[3], const int8_t* data, const int32_t dim[3], int n_it, double res, int conv,
                         const double origin[3], int use_new, double* points, double* normals, int8_t* data_out) {
  const size_t n = (size_t)dim[0] * dim[1] * dim[2];
  std::vector<convex_decomp_lib::data_type> grid(data, data + n);
  const Vec3i s(seed[0], seed[1], seed[2]), d(dim[0], dim[1], dim[2]);
  const Vec3f o(origin[0], origin[1], origin[2]);
  const Polyhedron3D poly = use_new ? convex_decomp_lib::GetPolyOcta3DNew(s, grid, d, n_it, res, conv, o)
                                    : convex_decomp_lib::GetPolyOcta3D(s, grid, d, n_it, res, conv, o);
  const auto pn = poly.cal_normals();
  for (size_t i = 0; i < pn.size(); ++i)
    for (int a = 0; a < 3; ++a) points[3 * i + a] = pn[i].first(a), normals[3 * i + a] = pn[i].second(a);
  if (data_out) memcpy(data_out, grid.data(), n);
  return (int)pn.size();
}
}
