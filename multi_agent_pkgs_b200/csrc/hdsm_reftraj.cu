// Reference-trajectory generation for sm_100a: one warp per agent, SURVEY.md section 8(f) row 2.
//
// Replaces, per agent and per replanning step (multi_agent_planner/src/agent_class.cpp):
//   Agent::GenerateReferenceTrajectory   :1449-1553   start point on the path, sampling, velocity reference
//   Agent::SamplePath                    :1591-1663
//   Agent::KeepOnlyFreeReference         :1665-1693
//   Agent::ComputePathVelocity           :1695-1803   ray casts through the potential field + neighbour sweep
//   Agent::GetVelocityLimit              :1805-1817
//   voxel_grid_util::Raycast             voxel_grid_util/src/raycast.cpp:21-186 (via path_finding_util::IsLineClear)
// and writes `ref` in the layout hdsm_solve_batch_device reads ([n][N][6]) next to the full [n][N+1][6]
// trajectory the next step starts from.
//
// Design.  The path is a handful of segments; each is traversed voxel by voxel (Amanatides-Woo) by the
// whole warp in lock step - first only to learn whether it is clear, then, if so, again with the visited
// points dealt round-robin to the lanes so that the expensive part (one pow and one exp per visited voxel)
// runs 32 wide.  The neighbour sweep (every plan step against every other agent's plan, the O(N n_rob)
// part that needs the all-gathered table) puts neighbours on lanes.  Minima are order independent, the
// sampling arithmetic uses explicit round-to-nearest intrinsics in the reference's evaluation order; only
// pow / exp differ from the CPU's libm (last-bit differences of the velocity, parity to 1e-12).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <new>
#include <string>

#include "../../include/hdsm.h"

namespace hdsm_rt {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kOcc = 100, kUnknown = -1;
constexpr int kMaxPath = 32;  // way-points of path_samp held in shared memory

struct Args {
  hdsm_reftraj_params prm;
  int n, n_rob;
  const int8_t* grids;
  size_t grid_stride;
  const int32_t *grid_index, *dims, *n_path, *global_id, *nbr_begin, *nbr_end;
  const double *origins, *path, *prev_ref, *traj, *all_pos;
  const uint8_t *have_prev, *increment, *all_valid;
  double *ref, *ref_solver, *path_vel;
};

__device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double dvd(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ double norm3(double x, double y, double z) { return __dsqrt_rn(add(add(mul(x, x), mul(y, y)), mul(z, z))); }
__device__ __forceinline__ double warp_min(double v) {
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(kFull, v, o));
  return v;
}

struct Grid {
  const int8_t* data;
  int dim[3];
  __device__ __forceinline__ bool inside(int x, int y, int z) const { return x >= 0 && y >= 0 && z >= 0 && x < dim[0] && y < dim[1] && z < dim[2]; }
  __device__ __forceinline__ int get(int x, int y, int z) const {  // GetVoxelInt: -1 outside (voxel_grid.cpp:110-117)
    return inside(x, y, z) ? (int)__ldg(data + x + (size_t)y * dim[0] + (size_t)z * dim[0] * dim[1]) : -1;
  }
};

// GetVelocityLimit (:1805-1817)
__device__ __forceinline__ double velocity_limit(const hdsm_reftraj_params& P, double occ, double dist) {
  occ = fmin(fmax(occ, 0.0), 100.0);
  const double alpha = sub(1.0, mul(pow(dvd(occ, 100.0), P.sens_pot), dvd(1.0, exp(mul(P.sens_dist, dist)))));
  return add(P.path_vel_min, mul(sub(P.path_vel_max, P.path_vel_min), alpha));
}

// One Amanatides-Woo traversal (raycast.cpp:21-186), executed by every lane in lock step.  visit(k, x, y, z) is
// called for the k-th element of the reference's output vector; returns true when a collision ended the ray
// (col = collision_pt).  A ray of more than 1500 voxels (the reference throws) reports *too_long.
template <class F>
__device__ bool raycast(const Grid& G, const double s[3], const double e[3], double max_dist, double col[3], bool* too_long, F&& visit) {
  int c[3], st[3];
  double d[3], tm[3], td[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    c[a] = (int)floor(s[a]);
    const int ea = (int)floor(e[a]);
    d[a] = sub(e[a], s[a]);
    st[a] = ea == c[a] ? 0 : (ea < c[a] ? -1 : 1);
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    // intbound (raycast.cpp:10-19): smallest positive t with s + t d integer
    double ss = s[a], ds = d[a];
    if (ds < 0) ss = -ss, ds = -ds;
    ss = fmod(add(fmod(ss, 1.0), 1.0), 1.0);
    tm[a] = dvd(sub(1.0, ss), ds);
    td[a] = dvd((double)st[a], d[a]);
  }
  int n = 0;
  *too_long = false;
  if (st[0] == 0 && st[1] == 0 && st[2] == 0) {  // same voxel: (end, start) (:89-93)
    visit(0, e[0], e[1], e[2]);
    visit(1, s[0], s[1], s[2]);
    return false;
  }
  const double max2 = mul(max_dist, max_dist);
  double tmax = 0;
#pragma unroll 1
  for (;;) {
    const double t = fmin(1.0, tmax);
    const double rx = add(s[0], mul(t, d[0])), ry = add(s[1], mul(t, d[1])), rz = add(s[2], mul(t, d[2]));
    if (G.inside(c[0], c[1], c[2])) {
      if (G.get(c[0], c[1], c[2]) == kOcc && tmax <= 1) {
        col[0] = rx, col[1] = ry, col[2] = rz;
        return true;
      }
      visit(n, rx, ry, rz);
      ++n;
      const double dx = sub((double)c[0], s[0]), dy = sub((double)c[1], s[1]), dz = sub((double)c[2], s[2]);
      if (add(add(mul(dx, dx), mul(dy, dy)), mul(dz, dz)) > max2) break;
      if (n > 1500) {
        *too_long = true;
        break;
      }
    }
    if (tmax >= 1) break;
    int ax;
    if ((tm[0] < tm[1] && st[0] != 0) || st[1] == 0) ax = ((tm[0] < tm[2] && st[0] != 0) || st[2] == 0) ? 0 : 2;
    else ax = ((tm[1] < tm[2] && st[1] != 0) || st[2] == 0) ? 1 : 2;
    // dynamic component select without local-memory arrays
    tmax = ax == 0 ? tm[0] : ax == 1 ? tm[1] : tm[2];
    if (ax == 0) c[0] += st[0], tm[0] = add(tm[0], td[0]);
    else if (ax == 1) c[1] += st[1], tm[1] = add(tm[1], td[1]);
    else c[2] += st[2], tm[2] = add(tm[2], td[2]);
  }
  return false;
}

__global__ void __launch_bounds__(32) reftraj_kernel(const Args A) {
  __shared__ double ps[kMaxPath + 1][3];  // path_samp
  __shared__ double pts[HDSM_MAX_HOR + 2][3];
  const int agent = blockIdx.x, lane = threadIdx.x;
  if (agent >= A.n) return;
  const hdsm_reftraj_params& P = A.prm;
  const int N = P.n_hor, N1 = N + 1;
  Grid G;
  G.data = A.grids + (size_t)(A.grid_index ? A.grid_index[agent] : agent) * A.grid_stride;
  double org[3];
  for (int a = 0; a < 3; ++a) G.dim[a] = A.dims[3 * agent + a], org[a] = A.origins[3 * agent + a];
  const double* path = A.path + (size_t)agent * P.max_path * 3;
  const int n_path = min(A.n_path[agent], kMaxPath);

  // ---- starting point and the part of the path ahead of it (:1456-1495)
  double start[3];
  if (A.have_prev[agent]) {
    const double* sp = A.prev_ref + (size_t)agent * N1 * 3 + (A.increment[agent] ? 3 : 0);
    start[0] = sp[0], start[1] = sp[1], start[2] = sp[2];
  } else {
    start[0] = path[0], start[1] = path[1], start[2] = path[2];
  }
  int start_idx = 0;
  {
    bool on = false;
    if (lane < n_path - 1) {  // IsOnSegment (:1864-1884), one segment per lane, the first hit wins
      const double* s1 = path + 3 * lane;
      const double* s2 = s1 + 3;
      const auto dist = [](const double* a, const double* b) {
        const double x = sub(a[0], b[0]), y = sub(a[1], b[1]), z = sub(a[2], b[2]);
        return __dsqrt_rn(add(add(mul(x, x), mul(y, y)), mul(z, z)));
      };
      if (fabs(sub(add(dist(start, s1), dist(start, s2)), dist(s1, s2))) < 1e-6) {
        const double dot = add(add(mul(sub(start[0], s1[0]), sub(start[0], s2[0])), mul(sub(start[1], s1[1]), sub(start[1], s2[1]))),
                               mul(sub(start[2], s1[2]), sub(start[2], s2[2])));
        on = dot <= 0;
      }
    }
    const unsigned m = __ballot_sync(kFull, on);
    if (m) start_idx = __ffs(m);  // i + 1
  }
  const int ns = 1 + max(0, n_path - start_idx);
  if (lane < 3) ps[0][lane] = start[lane];
  for (int i = lane; i < (ns - 1) * 3; i += 32) ps[1 + i / 3][i % 3] = path[3 * start_idx + i];
  __syncwarp();

  // ---- ComputePathVelocity (:1695-1803)
  double vel = P.path_vel_max;
  if (ns >= 2) {
    double vmin = P.path_vel_max;  // per lane, reduced at the end
#pragma unroll 1
    for (int i = 0; i < ns - 1; ++i) {
      double s[3], e[3], col[3] = {-1, -1, -1};
      for (int a = 0; a < 3; ++a) {  // GetCoordLocal (voxel_grid.cpp:137-142)
        s[a] = dvd(sub(ps[i][a], org[a]), P.voxel_size);
        e[a] = dvd(sub(ps[i + 1][a], org[a]), P.voxel_size);
      }
      const double maxd = norm3(sub(s[0], e[0]), sub(s[1], e[1]), sub(s[2], e[2]));
      bool too_long;
      const bool hit = raycast(G, s, e, maxd, col, &too_long, [](int, double, double, double) {});
      if (too_long) break;
      if (!hit) {  // clear: every visited point and the start limit the speed (:1721-1748)
        int total = 0;
        const auto visit = [&](int k, double x, double y, double z) {
          if ((k & 31) == lane) {
            double val = (double)G.get((int)x, (int)y, (int)z);
            if (val == -1) val = 100;
            // world-frame path start minus local-frame point, times the voxel size, as in the reference (:1735)
            const double dist = mul(norm3(sub(ps[0][0], x), sub(ps[0][1], y), sub(ps[0][2], z)), P.voxel_size);
            vmin = fmin(vmin, velocity_limit(P, val, dist));
          }
          total = k + 1;
        };
        raycast(G, s, e, maxd, col, &too_long, visit);
        visit(total, s[0], s[1], s[2]);  // visited_points.push_back(start)
      } else {  // collision: its voxel and its distance in voxel units (:1749-1766), then stop
        const int val = (int)(signed char)G.get((int)col[0], (int)col[1], (int)col[2]);
        vmin = fmin(vmin, velocity_limit(P, (double)val, norm3(sub(s[0], col[0]), sub(s[1], col[1]), sub(s[2], col[2]))));
        break;
      }
    }
    // other agents as obstacles whose weight decays along the horizon (:1769-1801): neighbours on lanes
    const int self = A.global_id[agent];
    const int nb0 = A.nbr_begin ? A.nbr_begin[agent] : 0, nb1 = A.nbr_end ? A.nbr_end[agent] : A.n_rob;
    const double* traj = A.traj + (size_t)agent * P.n_traj * 3;
#pragma unroll 1
    for (int i = 0; i < P.n_traj; ++i) {
      const double mx = traj[3 * i], my = traj[3 * i + 1], mz = traj[3 * i + 2];
      const double occ = mul(100.0, pow(P.sens_other_agents, (double)i));
      for (int j = nb0 + lane; j < nb1; j += 32) {
        if (j == self || !A.all_valid[j]) continue;
        const double* o = A.all_pos + ((size_t)j * P.n_traj + i) * 3;
        vmin = fmin(vmin, velocity_limit(P, occ, norm3(sub(mx, o[0]), sub(my, o[1]), sub(mz, o[2]))));
      }
    }
    vel = warp_min(vmin);
  }

  // ---- SamplePath (:1591-1663), uniform over the warp
  int np = 0;
  if (ns < 2) {
    for (int i = lane; i < N * 3; i += 32) pts[i / 3][i % 3] = ps[0][i % 3];
    np = N;
  } else {
    const double samp = mul(vel, P.dt);
    int idx = 1, ri = 0;
    double cur[3] = {ps[0][0], ps[0][1], ps[0][2]}, limit = samp;
    if (lane < 3) pts[0][lane] = cur[lane];
    np = 1;
#pragma unroll 1
    while (ri < N) {
      const double df[3] = {sub(ps[idx][0], cur[0]), sub(ps[idx][1], cur[1]), sub(ps[idx][2], cur[2])};
      const double dn = __dsqrt_rn(add(add(mul(df[0], df[0]), mul(df[1], df[1])), mul(df[2], df[2])));
      if (dn > limit) {
        for (int a = 0; a < 3; ++a) cur[a] = add(cur[a], dvd(mul(limit, df[a]), dn));
        if (lane < 3) pts[np][lane] = cur[lane];
        ++np, ++ri;
        limit = fmax(0.0, sub(samp, mul(P.path_vel_dec, P.dt)));
      } else {
        cur[0] = ps[idx][0], cur[1] = ps[idx][1], cur[2] = ps[idx][2];
        if (++idx == ns) {
          for (int i = ri; i < N; ++i, ++np)
            if (lane < 3) pts[np][lane] = ps[ns - 1][lane];
          break;
        }
        limit = sub(limit, dn);
      }
    }
  }
  __syncwarp();
  // ---- KeepOnlyFreeReference (:1665-1693): from the first unknown / occupied sample on, repeat the last free one
  {
    bool bad = false;
    if (lane >= 1 && lane < np) {
      const int v = G.get((int)dvd(sub(pts[lane][0], org[0]), P.voxel_size), (int)dvd(sub(pts[lane][1], org[1]), P.voxel_size),
                          (int)dvd(sub(pts[lane][2], org[2]), P.voxel_size));
      bad = v == kUnknown || v == kOcc;
    }
    const unsigned m = __ballot_sync(kFull, bad);
    if (m) {
      const int first = __ffs(m) - 1;
      double keep[3] = {pts[first - 1][0], pts[first - 1][1], pts[first - 1][2]};
      __syncwarp();
      if (lane >= first && lane < np)
        for (int a = 0; a < 3; ++a) pts[lane][a] = keep[a];
    }
  }
  __syncwarp();
  // ---- outputs: positions, velocity reference pointing backwards along the path (:1528-1546)
  double* ref = A.ref + (size_t)agent * N1 * 6;
  double v3[3] = {0, 0, 0};
  if (lane < N1) {
    const bool have = lane < np;
    if (have && np > 1) {
      const int i = min(lane, np - 2);  // the last point repeats the velocity of the one before
      const double x = sub(pts[i][0], pts[i + 1][0]), y = sub(pts[i][1], pts[i + 1][1]), z = sub(pts[i][2], pts[i + 1][2]);
      const double dd = __dsqrt_rn(add(add(mul(x, x), mul(y, y)), mul(z, z)));
      if (dd > 1e-2) v3[0] = dvd(mul(vel, x), dd), v3[1] = dvd(mul(vel, y), dd), v3[2] = dvd(mul(vel, z), dd);
    }
    for (int a = 0; a < 3; ++a) {
      const double p = have ? pts[lane][a] : 0.0, v = have ? v3[a] : 0.0;
      ref[6 * lane + a] = p, ref[6 * lane + 3 + a] = v;
      if (A.ref_solver && lane < N) A.ref_solver[((size_t)agent * N + lane) * 6 + a] = p, A.ref_solver[((size_t)agent * N + lane) * 6 + 3 + a] = v;
    }
  }
  if (lane == 0) A.path_vel[agent] = vel;
}

}  // namespace hdsm_rt

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
struct hdsm_reftraj {
  hdsm_reftraj_params prm{};
  int device = 0, max_agents = 0, max_grids = 0, max_rob = 0;
  size_t grid_stride = 0;
  cudaStream_t stream = nullptr;
  unsigned char *d_in = nullptr, *d_out = nullptr, *h_out = nullptr;
  size_t in_cap = 0, out_cap = 0;
  int64_t launches = 0;
  std::string err;
};

namespace {
int rfail(hdsm_reftraj* h, int code, const std::string& msg) {
  if (h) h->err = msg;
  return code;
}
#define RCU(call)                                                                              \
  do {                                                                                         \
    cudaError_t e_ = (call);                                                                   \
    if (e_ != cudaSuccess) return rfail(h, HDSM_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
  } while (0)
size_t ral256(size_t x) { return (x + 255) & ~size_t(255); }
}  // namespace

extern "C" {

int hdsm_reftraj_create(const hdsm_reftraj_params* p, int max_agents, int max_grids, size_t grid_stride, int device,
                        hdsm_reftraj** out) {
  if (!p || !out || max_agents < 1 || max_grids < 1 || grid_stride < 1) return HDSM_ERR_INVALID;
  if (p->n_hor < 1 || p->n_hor > HDSM_MAX_HOR || p->max_path < 1 || p->max_path > hdsm_rt::kMaxPath || p->n_traj < 0 ||
      !(p->dt > 0) || !(p->voxel_size > 0) || !(p->path_vel_max >= p->path_vel_min))
    return HDSM_ERR_INVALID;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return HDSM_ERR_CUDA;  // no CPU fallback
  hdsm_reftraj* h = new (std::nothrow) hdsm_reftraj();
  if (!h) return HDSM_ERR_INVALID;
  h->prm = *p, h->device = device, h->max_agents = max_agents, h->max_grids = max_grids, h->grid_stride = grid_stride;
  if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete h;
    return HDSM_ERR_CUDA;
  }
  *out = h;
  return HDSM_OK;
}

void hdsm_reftraj_destroy(hdsm_reftraj* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->d_in) cudaFree(h->d_in);
  if (h->d_out) cudaFree(h->d_out);
  if (h->h_out) cudaFreeHost(h->h_out);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

const char* hdsm_reftraj_last_error(const hdsm_reftraj* h) { return h ? h->err.c_str() : "null handle"; }
int64_t hdsm_reftraj_launch_count(const hdsm_reftraj* h) { return h ? h->launches : 0; }

int hdsm_reftraj_batch_device(hdsm_reftraj* h, int n, const int8_t* grids, const int32_t* grid_index, const int32_t* dims,
                              const double* origins, const double* path, const int32_t* n_path, const double* prev_ref,
                              const uint8_t* have_prev, const uint8_t* increment, const double* traj,
                              const int32_t* global_id, const int32_t* nbr_begin, const int32_t* nbr_end,
                              const double* all_pos, const uint8_t* all_valid, int n_rob, double* ref, double* ref_solver,
                              double* path_vel, void* stream) {
  if (!h) return HDSM_ERR_INVALID;
  if (n < 0 || n_rob < 0 || !grids || !dims || !origins || !path || !n_path || !prev_ref || !have_prev || !increment ||
      !global_id || !ref || !path_vel || ((nbr_begin == nullptr) != (nbr_end == nullptr)) ||
      (h->prm.n_traj > 0 && n_rob > 0 && (!traj || !all_pos || !all_valid)))
    return rfail(h, HDSM_ERR_INVALID, "null or inconsistent argument");
  if (n > h->max_agents) return rfail(h, HDSM_ERR_CAPACITY, "n exceeds max_agents");
  if (n == 0) return HDSM_OK;
  RCU(cudaSetDevice(h->device));
  hdsm_rt::Args a{};
  a.prm = h->prm, a.n = n, a.n_rob = n_rob, a.grids = grids, a.grid_stride = h->grid_stride, a.grid_index = grid_index;
  a.dims = dims, a.n_path = n_path, a.global_id = global_id, a.nbr_begin = nbr_begin, a.nbr_end = nbr_end;
  a.origins = origins, a.path = path, a.prev_ref = prev_ref, a.traj = traj, a.all_pos = all_pos;
  a.have_prev = have_prev, a.increment = increment, a.all_valid = all_valid;
  a.ref = ref, a.ref_solver = ref_solver, a.path_vel = path_vel;
  if (n_rob == 0) a.prm.n_traj = 0;  // nobody to sweep against
  cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : h->stream;
  hdsm_rt::reftraj_kernel<<<n, 32, 0, s>>>(a);
  h->launches += 1;
  RCU(cudaGetLastError());
  return HDSM_OK;
}

int hdsm_reftraj_batch(hdsm_reftraj* h, int n, int n_grids, const int8_t* grids, const int32_t* grid_index, const int32_t* dims,
                       const double* origins, const double* path, const int32_t* n_path, const double* prev_ref,
                       const uint8_t* have_prev, const uint8_t* increment, const double* traj, const int32_t* global_id,
                       const int32_t* nbr_begin, const int32_t* nbr_end, const double* all_pos, const uint8_t* all_valid,
                       int n_rob, double* ref, double* path_vel) {
  if (!h) return HDSM_ERR_INVALID;
  if (n < 0 || n_grids < 1 || n_rob < 0 || !grids || !dims || !origins || !path || !n_path || !prev_ref || !have_prev ||
      !increment || !global_id || !ref || !path_vel)
    return rfail(h, HDSM_ERR_INVALID, "null argument");
  if (n > h->max_agents || n_grids > h->max_grids) return rfail(h, HDSM_ERR_CAPACITY, "n / n_grids exceed the handle's capacity");
  if (!grid_index && n_grids < n) return rfail(h, HDSM_ERR_INVALID, "grid_index is required when agents share grids");
  if (n == 0) return HDSM_OK;
  for (int i = 0; i < n; ++i) {
    const int gi = grid_index ? grid_index[i] : i;
    if (gi < 0 || gi >= n_grids) return rfail(h, HDSM_ERR_INVALID, "grid_index out of range");
    const int32_t* d = dims + 3 * i;
    if (d[0] < 1 || d[1] < 1 || d[2] < 1 || (size_t)d[0] * d[1] * d[2] > h->grid_stride)
      return rfail(h, HDSM_ERR_INVALID, "grid dimensions exceed grid_stride");
    if (n_path[i] < 1 || n_path[i] > h->prm.max_path) return rfail(h, HDSM_ERR_INVALID, "n_path outside 1..max_path");
    if (global_id[i] < 0 || (n_rob > 0 && global_id[i] >= n_rob)) return rfail(h, HDSM_ERR_INVALID, "global_id out of range");
  }
  RCU(cudaSetDevice(h->device));
  const size_t N = (size_t)n, N1 = (size_t)h->prm.n_hor + 1, NT = (size_t)h->prm.n_traj;
  struct Seg {
    const void* src;
    size_t bytes, off;
  };
  Seg in[16];
  int ni = 0;
  size_t off = 0;
  const auto put = [&](const void* p, size_t bytes) {
    in[ni] = Seg{p, (p && bytes) ? bytes : 0, off};
    if (p && bytes) off += ral256(bytes);
    return ni++;
  };
  const int i_grid = put(grids, (size_t)n_grids * h->grid_stride), i_gi = put(grid_index, N * 4), i_dim = put(dims, N * 12);
  const int i_org = put(origins, N * 24), i_path = put(path, N * h->prm.max_path * 24), i_np = put(n_path, N * 4);
  const int i_pr = put(prev_ref, N * N1 * 24), i_hp = put(have_prev, N), i_inc = put(increment, N), i_tr = put(traj, N * NT * 24);
  const int i_id = put(global_id, N * 4), i_b = put(nbr_begin, N * 4), i_e = put(nbr_end, N * 4);
  const int i_ap = put(all_pos, (size_t)n_rob * NT * 24), i_av = put(all_valid, (size_t)n_rob);
  if (off > h->in_cap) {
    if (h->d_in) cudaFree(h->d_in);
    h->d_in = nullptr, h->in_cap = 0;
    RCU(cudaMalloc(&h->d_in, off));
    h->in_cap = off;
  }
  const size_t o_ref = 0, o_vel = ral256(N * N1 * 48), out_bytes = o_vel + ral256(N * 8);
  if (out_bytes > h->out_cap) {
    if (h->d_out) cudaFree(h->d_out);
    if (h->h_out) cudaFreeHost(h->h_out);
    h->d_out = h->h_out = nullptr, h->out_cap = 0;
    RCU(cudaMalloc(&h->d_out, out_bytes));
    RCU(cudaMallocHost(&h->h_out, out_bytes));
    h->out_cap = out_bytes;
  }
  for (int k = 0; k < ni; ++k)
    if (in[k].bytes) RCU(cudaMemcpyAsync(h->d_in + in[k].off, in[k].src, in[k].bytes, cudaMemcpyHostToDevice, h->stream));
  const auto dp = [&](int k) -> const void* { return in[k].bytes ? h->d_in + in[k].off : nullptr; };
  const int rc = hdsm_reftraj_batch_device(
      h, n, (const int8_t*)dp(i_grid), (const int32_t*)dp(i_gi), (const int32_t*)dp(i_dim), (const double*)dp(i_org),
      (const double*)dp(i_path), (const int32_t*)dp(i_np), (const double*)dp(i_pr), (const uint8_t*)dp(i_hp),
      (const uint8_t*)dp(i_inc), (const double*)dp(i_tr), (const int32_t*)dp(i_id), (const int32_t*)dp(i_b),
      (const int32_t*)dp(i_e), (const double*)dp(i_ap), (const uint8_t*)dp(i_av), n_rob, (double*)(h->d_out + o_ref), nullptr,
      (double*)(h->d_out + o_vel), h->stream);
  if (rc != HDSM_OK) return rc;
  RCU(cudaMemcpyAsync(h->h_out, h->d_out, out_bytes, cudaMemcpyDeviceToHost, h->stream));
  RCU(cudaStreamSynchronize(h->stream));
  memcpy(ref, h->h_out + o_ref, N * N1 * 48);
  memcpy(path_vel, h->h_out + o_vel, N * 8);
  return HDSM_OK;
}

}  // extern "C"
