// Per-ray and per-voxel arithmetic of the local-map acquisition kernel (csrc/hdsm_sense.cu), written once for
// the device and for the host so that the tests can run the very same code on the CPU in a scrambled ray order
// and compare it with the sequential restatement of the reference (the product only ever runs the device build).
//
// Follows mapping_util/src/map_builder.cpp: frame :89-118, ClearLine :367-432 (on top of
// voxel_grid_util::Raycast, voxel_grid_util/src/raycast.cpp:21-186), RaycastAndClear ray order :280-329,
// MergeVoxelGrids :242-278, ClearVoxelsCenter :434-447.
//
// Order independence.  The reference casts ~14 000 rays one after the other into one output grid; each ray writes
// "occupied" to the voxel behind its collision point and then "free" to every voxel it crossed, and a later write
// replaces an earlier one.  Here every write carries the key 2 (seq + 1) + is_free, seq = the ray's position in the
// reference's loop order, and a voxel keeps the largest key it has seen (atomicMax on the device): the voxel's
// final value is that of the write the reference would have made last, whatever order the rays run in.
//
// All double arithmetic uses explicit round-to-nearest operations on the device (no FMA contraction); host
// translation units including this header must be compiled with -ffp-contract=off.
#ifndef HDSM_SENSE_CORE_H_
#define HDSM_SENSE_CORE_H_
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define SN_HD __host__ __device__ __forceinline__
#else
#define SN_HD static inline
#endif
#if defined(__CUDA_ARCH__)
#define SN_MUL(a, b) __dmul_rn((a), (b))
#define SN_ADD(a, b) __dadd_rn((a), (b))
#define SN_SUB(a, b) __dsub_rn((a), (b))
#define SN_DIV(a, b) __ddiv_rn((a), (b))
#define SN_SQRT(a) __dsqrt_rn((a))
#else
#define SN_MUL(a, b) ((a) * (b))
#define SN_ADD(a, b) ((a) + (b))
#define SN_SUB(a, b) ((a) - (b))
#define SN_DIV(a, b) ((a) / (b))
#define SN_SQRT(a) sqrt((a))
#endif

namespace hdsm_sn {

struct Frame {
  double origin[3];     // origin of the local grid (world)
  double pos_local[3];  // the agent in local voxel coordinates (VoxelGrid::GetCoordLocal)
  int dim[3];           // floor(range / voxel)
  int start[3];         // index of the local grid's first voxel in the environment grid
};

// map_builder.cpp:89-118 and :170
SN_HD void make_frame(double voxel, const double range[3], const double origin_env[3], const double pos[3], Frame& F) {
  for (int a = 0; a < 3; ++a) {
    double o = SN_SUB(pos[a], SN_DIV(range[a], 2.0));
    o = SN_ADD(SN_MUL(round(SN_DIV(SN_SUB(o, origin_env[a]), voxel)), voxel), origin_env[a]);
    F.origin[a] = o;
    F.dim[a] = (int)floor(SN_DIV(range[a], voxel));
    F.start[a] = (int)round(SN_DIV(SN_SUB(o, origin_env[a]), voxel));
    F.pos_local[a] = SN_DIV(SN_SUB(pos[a], o), voxel);
  }
}

SN_HD bool inside(const int dim[3], int x, int y, int z) { return x >= 0 && y >= 0 && z >= 0 && x < dim[0] && y < dim[1] && z < dim[2]; }

SN_HD int ray_count(const int dim[3]) { return 2 * (dim[0] * dim[1] + dim[0] * dim[2] + dim[1] * dim[2]); }

// end point of ray `seq` in the order of RaycastAndClear (:295-325): floor and ceiling, y walls, x walls
SN_HD void ray_end(const int dim[3], int seq, double end[3]) {
  const int side = seq & 1;
  int q = seq >> 1, i, j, k;
  const int n_fc = dim[0] * dim[1], n_y = dim[0] * dim[2];
  if (q < n_fc) {
    i = q / dim[1], j = q - i * dim[1], k = side ? dim[2] - 1 : 0;
  } else if (q < n_fc + n_y) {
    q -= n_fc;
    i = q / dim[2], k = q - i * dim[2], j = side ? dim[1] - 1 : 0;
  } else {
    q -= n_fc + n_y;
    j = q / dim[2], k = q - j * dim[2], i = side ? dim[0] - 1 : 0;
  }
  end[0] = i + 0.5, end[1] = j + 0.5, end[2] = k + 0.5;
}

SN_HD double dot3(const double* a, const double* b) { return SN_ADD(SN_ADD(SN_MUL(a[0], b[0]), SN_MUL(a[1], b[1])), SN_MUL(a[2], b[2])); }
SN_HD void normalize3(double* v) {
  const double z = dot3(v, v);
  if (z > 0) {
    const double n = SN_SQRT(z);
    v[0] = SN_DIV(v[0], n), v[1] = SN_DIV(v[1], n), v[2] = SN_DIV(v[2], n);
  }
}

// field-of-view test of ClearLine (:375-393); rot = rot_mat_cam_ row-major, columns are the body axes
SN_HD bool in_fov(const double* rot, const double s[3], const double e[3], double cos_half_x, double cos_half_y) {
  const double xb[3] = {rot[0], rot[3], rot[6]}, yb[3] = {rot[1], rot[4], rot[7]}, zb[3] = {rot[2], rot[5], rot[8]};
  const double dir[3] = {SN_SUB(e[0], s[0]), SN_SUB(e[1], s[1]), SN_SUB(e[2], s[2])};
  const double dy = dot3(dir, yb), dz = dot3(dir, zb);
  double dir_xz[3], dir_xy[3];
  for (int a = 0; a < 3; ++a) dir_xz[a] = SN_SUB(dir[a], SN_MUL(dy, yb[a])), dir_xy[a] = SN_SUB(dir[a], SN_MUL(dz, yb[a]));
  normalize3(dir_xz);
  normalize3(dir_xz);  // as written in the reference (:386): dir_xz twice, dir_xy never
  return dot3(dir_xz, xb) > cos_half_y && dot3(dir_xy, xb) > cos_half_x;
}

SN_HD uint32_t key_free(int seq) { return 2u * (uint32_t)(seq + 1) + 1u; }
SN_HD uint32_t key_occ(int seq) { return 2u * (uint32_t)(seq + 1); }
SN_HD int8_t key_value(uint32_t key) { return key == 0 ? (int8_t)-1 : ((key & 1u) ? (int8_t)0 : (int8_t)100); }

// fmod(x, 1.0): x minus its integer part, exact in binary floating point (the library fmod is a long loop on the device;
// a -0.0 result comes out as +0.0, which the + 1 that follows in intbound makes indistinguishable)
SN_HD double frac1(double x) { return SN_SUB(x, trunc(x)); }

// ClearLine (:395-431) for one ray.  occ(x, y, z): is the (inside) voxel occupied in the cropped grid;
// put(cell, key): offer `key` to voxel `cell` (x + y dx + z dx dy).
template <class Occ, class Put>
SN_HD void clear_line(const int dim[3], const double s[3], const double e[3], int seq, Occ&& occ, Put&& put) {
  const uint32_t kf = key_free(seq);
  int c[3], st[3];
  double d[3], tm[3], td[3];
  for (int a = 0; a < 3; ++a) {
    c[a] = (int)floor(s[a]);
    const int ea = (int)floor(e[a]);
    d[a] = SN_SUB(e[a], s[a]);
    st[a] = ea == c[a] ? 0 : (ea < c[a] ? -1 : 1);
  }
  for (int a = 0; a < 3; ++a) {  // intbound (raycast.cpp:10-19)
    double ss = s[a], ds = d[a];
    if (ds < 0) ss = -ss, ds = -ds;
    ss = frac1(SN_ADD(frac1(ss), 1.0));  // mod(s, 1) = fmod(fmod(s, 1) + 1, 1)
    tm[a] = SN_DIV(SN_SUB(1.0, ss), ds);
    td[a] = SN_DIV((double)st[a], d[a]);
  }
  double px = 0, py = 0, pz = 0;
  int n = 0;
  // the reference collects the visited points and then frees the voxel under the middle of every consecutive pair
  const auto emit = [&](double x, double y, double z) {
    if (n > 0) {
      // (a + b) / 2.0 as a multiplication by 0.5: the same correctly rounded value, without the device's division routine
      const int vx = (int)SN_MUL(SN_ADD(px, x), 0.5), vy = (int)SN_MUL(SN_ADD(py, y), 0.5), vz = (int)SN_MUL(SN_ADD(pz, z), 0.5);
      if (inside(dim, vx, vy, vz)) put(vx + dim[0] * (vy + dim[1] * vz), kf);
    }
    px = x, py = y, pz = z, ++n;
  };
  if (st[0] == 0 && st[1] == 0 && st[2] == 0) {  // same voxel: Raycast returns (end, start) (:89-93); clear -> + end
    emit(e[0], e[1], e[2]);
    emit(s[0], s[1], s[2]);
    emit(e[0], e[1], e[2]);
    return;
  }
  const double ex = SN_SUB(s[0], e[0]), ey = SN_SUB(s[1], e[1]), ez = SN_SUB(s[2], e[2]);
  const double max_dist = SN_SQRT(SN_ADD(SN_ADD(SN_MUL(ex, ex), SN_MUL(ey, ey)), SN_MUL(ez, ez)));  // (start - end).norm() (:398)
  const double max2 = SN_MUL(max_dist, max_dist);
  double tmax = 0;
  for (;;) {
    const double t = tmax < 1.0 ? tmax : 1.0;
    const double rx = SN_ADD(s[0], SN_MUL(t, d[0])), ry = SN_ADD(s[1], SN_MUL(t, d[1])), rz = SN_ADD(s[2], SN_MUL(t, d[2]));
    if (inside(dim, c[0], c[1], c[2])) {
      if (tmax <= 1 && occ(c[0], c[1], c[2])) {
        // collision: the voxel just behind the collision point becomes occupied (:407-411), the collision point is
        // the last visited point and the end point is not appended
        const int lx = (int)SN_ADD(SN_MUL(d[0], 1e-7), rx), ly = (int)SN_ADD(SN_MUL(d[1], 1e-7), ry), lz = (int)SN_ADD(SN_MUL(d[2], 1e-7), rz);
        if (inside(dim, lx, ly, lz)) put(lx + dim[0] * (ly + dim[1] * lz), key_occ(seq));
        emit(rx, ry, rz);
        return;
      }
      emit(rx, ry, rz);
      const double dx = SN_SUB((double)c[0], s[0]), dy = SN_SUB((double)c[1], s[1]), dz = SN_SUB((double)c[2], s[2]);
      if (SN_ADD(SN_ADD(SN_MUL(dx, dx), SN_MUL(dy, dy)), SN_MUL(dz, dz)) > max2) break;
    }
    if (tmax >= 1) break;
    int ax;  // the neighbouring voxel whose boundary is crossed first (:160-181)
    if ((tm[0] < tm[1] && st[0] != 0) || st[1] == 0)
      ax = ((tm[0] < tm[2] && st[0] != 0) || st[2] == 0) ? 0 : 2;
    else
      ax = ((tm[1] < tm[2] && st[1] != 0) || st[2] == 0) ? 1 : 2;
    // dynamic indexing would put c / tm / td into local memory on the device
    if (ax == 0) tmax = tm[0], c[0] += st[0], tm[0] = SN_ADD(tm[0], td[0]);
    else if (ax == 1) tmax = tm[1], c[1] += st[1], tm[1] = SN_ADD(tm[1], td[1]);
    else tmax = tm[2], c[2] += st[2], tm[2] = SN_ADD(tm[2], td[2]);
  }
  emit(e[0], e[1], e[2]);  // line clear: visited_points.push_back(end) (:404-405)
}

// value of the kept grid (voxel_grid_curr_) at local cell (x, y, z) of the NEW grid, for MergeVoxelGrids (:242-278);
// without a kept grid: the all-unknown grid with the 5x5x5 cube around the agent freed (:160-167, :434-447)
SN_HD int8_t old_value(const int8_t* old_grid, bool have_old, const int dim[3], const int off[3], const int mid[3], int x, int y, int z) {
  if (have_old) {
    const int ox = x + off[0], oy = y + off[1], oz = z + off[2];
    return inside(dim, ox, oy, oz) ? old_grid[ox + dim[0] * (oy + dim[1] * oz)] : (int8_t)-1;
  }
  return (x >= mid[0] - 2 && x <= mid[0] + 2 && y >= mid[1] - 2 && y <= mid[1] + 2 && z >= mid[2] - 2 && z <= mid[2] + 2) ? (int8_t)0 : (int8_t)-1;
}

// offset of the new grid in the old one, in voxels (:255-258)
SN_HD void merge_offset(const double origin_new[3], const double origin_old[3], double voxel, int off[3]) {
  for (int a = 0; a < 3; ++a) off[a] = (int)round(SN_DIV(SN_SUB(origin_new[a], origin_old[a]), voxel));
}

// the cropped environment value (:120-153) for local cell (x, y, z)
SN_HD int8_t crop_value(const int8_t* env, const int dim_env[3], const int start[3], bool free_grid, int x, int y, int z) {
  const int ie = x + start[0], je = y + start[1], ke = z + start[2];
  int8_t v = inside(dim_env, ie, je, ke) ? env[(size_t)ie + (size_t)dim_env[0] * ((size_t)je + (size_t)dim_env[1] * ke)] : (int8_t)-1;
  if (free_grid) {
    if (v == -1) v = 0;
  } else if (v == 0) {
    v = -1;
  }
  return v;
}

}  // namespace hdsm_sn
#endif  // HDSM_SENSE_CORE_H_
