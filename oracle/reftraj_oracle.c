/* reftraj_oracle.c - CPU restatement of the reference-trajectory generator.
 *
 * TEST INFRASTRUCTURE ONLY: imported by tests/ and bench legs; the product never links or calls it.
 *
 * Restates, in plain C:
 *   rt_raycast          voxel_grid_util::Raycast               voxel_grid_util/src/raycast.cpp:21-186
 *                       (the core of path_finding_util::IsLineClear, path_tools.cpp:148-180)
 *   rt_velocity_limit   Agent::GetVelocityLimit                multi_agent_planner/src/agent_class.cpp:1805-1817
 *   rt_path_velocity    Agent::ComputePathVelocity             :1695-1803
 *   rt_generate         Agent::GenerateReferenceTrajectory     :1449-1553  (+ IsOnSegment :1864-1884,
 *                       SamplePath :1591-1663, KeepOnlyFreeReference :1665-1693)
 *
 * Parity: rt_raycast is PINNED - tests/test_reftraj_oracle.py compares its visited points and collision
 * point bit for bit with the reference's own Raycast (raycast.cpp and voxel_grid.cpp compiled unmodified
 * into oracle/_ref/libref_voxel.so).  The rest lives in agent_class.cpp and is PINNED too: that file compiles
 * unmodified on the stand-in ROS / Gurobi headers of oracle/ref_shim/ (oracle/_ref/libref_agent.so) and its own
 * GenerateReferenceTrajectory gives bit-identical references and path velocities (tests/test_ref_agent.py, fixture
 * tests/golden/reftraj_node_ref.npz).  The restatement keeps the reference's frame mix-up in ComputePathVelocity (the distance of a
 * visited voxel is taken between the path start in WORLD metres and the voxel in LOCAL voxel units, :1735)
 * and the collision distance left in voxel units (:1755).  pow / exp come from libm here and from CUDA's
 * math library in the kernel: parity of the velocity is to 1e-12 relative, not bit for bit.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define RT_OCC 100
#define RT_UNKNOWN (-1)
#define RT_MAX_VISITED 1502

typedef struct reftraj_params {
  int32_t n_hor;      /* n_hor_ */
  int32_t max_path;   /* row stride of the path array */
  int32_t n_traj;     /* points of traj_curr_ (N + 1, or 0 before the first plan) */
  int32_t reserved;
  double dt;
  double path_vel_min, path_vel_max, path_vel_dec;
  double sens_dist, sens_pot, sens_other_agents;
  double voxel;
} reftraj_params;

static int vg_inside(const int dim[3], int x, int y, int z) { return x >= 0 && y >= 0 && z >= 0 && x < dim[0] && y < dim[1] && z < dim[2]; }
static int vg_get(const int8_t* data, const int dim[3], int x, int y, int z) { /* GetVoxelInt: -1 outside (voxel_grid.cpp:110-117) */
  return vg_inside(dim, x, y, z) ? data[x + y * dim[0] + z * dim[0] * dim[1]] : -1;
}

static double rt_mod(double value, double modulus) { return fmod(fmod(value, modulus) + modulus, modulus); }
static double rt_intbound(double s, double ds) { /* smallest positive t with s + t ds integer (raycast.cpp:10-19) */
  if (ds < 0) return rt_intbound(-s, -ds);
  s = rt_mod(s, 1);
  return (1 - s) / ds;
}
static int rt_sign(int x) { return x == 0 ? 0 : x < 0 ? -1 : 1; }

/* Amanatides-Woo traversal from start to end (local voxel coordinates).  visited receives the reference's
 * output vector (up to cap points); collision = (-1,-1,-1) when the line is clear.  Returns the count. */
int rt_raycast(const int8_t* data, const int32_t dim_in[3], const double start[3], const double end[3], double max_dist,
               double* visited, int cap, double collision[3]) {
  const int dim[3] = {dim_in[0], dim_in[1], dim_in[2]};
  int c[3], e[3], step[3];
  double d[3], tmax3[3], tdelta[3];
  for (int a = 0; a < 3; ++a) {
    c[a] = (int)floor(start[a]), e[a] = (int)floor(end[a]);
    d[a] = end[a] - start[a];
    step[a] = rt_sign((int)(double)(e[a] - c[a]));
  }
  for (int a = 0; a < 3; ++a) tmax3[a] = rt_intbound(start[a], d[a]), tdelta[a] = (double)step[a] / d[a];
  collision[0] = collision[1] = collision[2] = -1;
  int n = 0;
#define PUSH(px, py, pz)                                                     \
  do {                                                                       \
    if (n < cap) visited[3 * n] = (px), visited[3 * n + 1] = (py), visited[3 * n + 2] = (pz); \
    ++n;                                                                     \
  } while (0)
  if (step[0] == 0 && step[1] == 0 && step[2] == 0) { /* same voxel: (end, start) (:89-93) */
    PUSH(end[0], end[1], end[2]);
    PUSH(start[0], start[1], start[2]);
    return n;
  }
  const double max2 = max_dist * max_dist;
  double tmax = 0;
  for (;;) {
    const double t = tmax < 1.0 ? tmax : 1.0;
    const double real[3] = {start[0] + t * d[0], start[1] + t * d[1], start[2] + t * d[2]};
    if (vg_inside(dim, c[0], c[1], c[2])) {
      if (vg_get(data, dim, c[0], c[1], c[2]) == RT_OCC && tmax <= 1) {
        collision[0] = real[0], collision[1] = real[1], collision[2] = real[2];
        PUSH(real[0], real[1], real[2]);
        break;
      }
      PUSH(real[0], real[1], real[2]);
      const double dx = c[0] - start[0], dy = c[1] - start[1], dz = c[2] - start[2];
      if ((dx * dx + dy * dy) + dz * dz > max2) break;
      if (n > 1500) return -1; /* the reference throws here */
    }
    if (tmax >= 1) break;
    /* step into the neighbouring voxel whose boundary is crossed first (:160-181) */
    int ax;
    if ((tmax3[0] < tmax3[1] && step[0] != 0) || step[1] == 0)
      ax = ((tmax3[0] < tmax3[2] && step[0] != 0) || step[2] == 0) ? 0 : 2;
    else
      ax = ((tmax3[1] < tmax3[2] && step[1] != 0) || step[2] == 0) ? 1 : 2;
    tmax = tmax3[ax];
    c[ax] += step[ax];
    tmax3[ax] += tdelta[ax];
  }
#undef PUSH
  return n;
}

double rt_velocity_limit(const reftraj_params* P, double occ_val, double dist_start) { /* :1805-1817 */
  if (occ_val < 0) occ_val = 0;
  if (occ_val > 100) occ_val = 100;
  const double alpha = (1 - pow(occ_val / 100, P->sens_pot) * (1 / exp(P->sens_dist * dist_start)));
  return P->path_vel_min + (P->path_vel_max - P->path_vel_min) * alpha;
}

static double norm3(double x, double y, double z) { return sqrt((x * x + y * y) + z * z); }

/* ComputePathVelocity (:1695-1803).  path [n_path][3] world metres (path[0] = the sampling start);
 * traj [n_traj][3] own plan positions; all_pos [n_rob][n_traj][3], all_valid [n_rob]; self = own id. */
double rt_path_velocity(const reftraj_params* P, const int8_t* data, const int32_t dim[3], const double origin[3],
                        const double* path, int n_path, const double* traj, const double* all_pos, const uint8_t* all_valid,
                        int nbr_begin, int nbr_end, int self) {
  double vel = P->path_vel_max;
  const int idim[3] = {dim[0], dim[1], dim[2]};
  double* visited = (double*)malloc(sizeof(double) * 3 * RT_MAX_VISITED);
  for (int i = 0; i < n_path - 1; ++i) {
    double s[3], e[3], col[3];
    for (int a = 0; a < 3; ++a) { /* GetCoordLocal: (p - origin) / vox (voxel_grid.cpp:137-142) */
      s[a] = (path[3 * i + a] - origin[a]) / P->voxel;
      e[a] = (path[3 * (i + 1) + a] - origin[a]) / P->voxel;
    }
    const double maxd = norm3(s[0] - e[0], s[1] - e[1], s[2] - e[2]);
    int nv = rt_raycast(data, dim, s, e, maxd, visited, RT_MAX_VISITED - 1, col);
    if (nv < 0) break;
    if (col[0] == -1) { /* clear: every visited voxel and the start limit the speed (:1721-1748) */
      if (nv > RT_MAX_VISITED - 1) nv = RT_MAX_VISITED - 1;
      visited[3 * nv] = s[0], visited[3 * nv + 1] = s[1], visited[3 * nv + 2] = s[2];
      ++nv;
      for (int k = 0; k < nv; ++k) {
        const double* pt = visited + 3 * k;
        double val = (double)vg_get(data, idim, (int)pt[0], (int)pt[1], (int)pt[2]);
        if (val == -1) val = 100;
        /* world-frame path start minus local-frame voxel, as in the reference (:1735) */
        const double dist = norm3(path[0] - pt[0], path[1] - pt[1], path[2] - pt[2]) * P->voxel;
        const double v = rt_velocity_limit(P, val, dist);
        if (v < vel) vel = v;
      }
    } else { /* collision: its voxel and its distance in voxel units (:1749-1766), then stop */
      const int8_t val = (int8_t)vg_get(data, idim, (int)col[0], (int)col[1], (int)col[2]);
      const double v = rt_velocity_limit(P, (double)val, norm3(s[0] - col[0], s[1] - col[1], s[2] - col[2]));
      if (v < vel) vel = v;
      break;
    }
  }
  free(visited);
  /* other agents as obstacles whose weight decays along the horizon (:1769-1801) */
  for (int i = 0; i < P->n_traj; ++i) {
    const double* me = traj + 3 * i;
    for (int j = nbr_begin; j < nbr_end; ++j) {
      if (j == self || !all_valid[j]) continue;
      const double* o = all_pos + ((size_t)j * P->n_traj + i) * 3;
      const double dist = norm3(me[0] - o[0], me[1] - o[1], me[2] - o[2]);
      const double v = rt_velocity_limit(P, 100 * pow(P->sens_other_agents, i), dist);
      if (v < vel) vel = v;
    }
  }
  return vel;
}

static double dist3(const double* a, const double* b) { /* sqrt(GetDistanceSquared) */
  const double x = a[0] - b[0], y = a[1] - b[1], z = a[2] - b[2];
  return sqrt(x * x + y * y + z * z);
}
static int on_segment(const double* pt, const double* s1, const double* s2) { /* IsOnSegment :1864-1884 */
  if (fabs(dist3(pt, s1) + dist3(pt, s2) - dist3(s1, s2)) < 1e-6) {
    const double a[3] = {pt[0] - s1[0], pt[1] - s1[1], pt[2] - s1[2]}, b[3] = {pt[0] - s2[0], pt[1] - s2[1], pt[2] - s2[2]};
    if (a[0] * b[0] + a[1] * b[1] + a[2] * b[2] <= 0) return 1;
  }
  return 0;
}

/* GenerateReferenceTrajectory for one agent.
 *   path [n_path][3]: path_curr_;  prev_ref [N+1][3] + have_prev: traj_ref_curr_ of the last step (positions);
 *   increment: increment_traj_ref_;  traj / all_pos / all_valid: plans for the neighbour sweep
 *   out: ref [N+1][6] (positions + velocity reference), *path_vel */
void rt_generate(const reftraj_params* P, const int8_t* data, const int32_t dim[3], const double origin[3], const double* path,
                 int n_path, const double* prev_ref, int have_prev, int increment, const double* traj, const double* all_pos,
                 const uint8_t* all_valid, int nbr_begin, int nbr_end, int self, double* ref, double* path_vel_out) {
  const int N = P->n_hor;
  const int idim[3] = {dim[0], dim[1], dim[2]};
  double* ps = (double*)malloc(sizeof(double) * 3 * (size_t)(n_path + 1)); /* path_samp */
  double start[3];
  if (have_prev) {
    const double* sp = prev_ref + (increment ? 3 : 0);
    start[0] = sp[0], start[1] = sp[1], start[2] = sp[2];
  } else {
    start[0] = path[0], start[1] = path[1], start[2] = path[2];
  }
  int start_idx = 0;
  for (int i = 0; i < n_path - 1; ++i)
    if (on_segment(start, path + 3 * i, path + 3 * (i + 1))) {
      start_idx = i + 1;
      break;
    }
  int ns = 0;
  memcpy(ps, start, sizeof start), ns = 1;
  for (int i = start_idx; i < n_path; ++i) memcpy(ps + 3 * ns++, path + 3 * i, sizeof(double) * 3);

  /* SamplePath (:1591-1663) */
  double* pts = (double*)calloc((size_t)(N + 2) * 3, sizeof(double));
  int np = 0;
  double vel = P->path_vel_max; /* path_vel_ keeps its previous value when the path has one point; max is the neutral choice */
  if (ns < 2) {
    for (int i = 0; i < N; ++i) memcpy(pts + 3 * np++, ps, sizeof(double) * 3);
  } else {
    vel = rt_path_velocity(P, data, dim, origin, ps, ns, traj, all_pos, all_valid, nbr_begin, nbr_end, self);
    const double samp = vel * P->dt;
    int idx = 1, ri = 0;
    double cur[3] = {ps[0], ps[1], ps[2]}, limit = samp;
    memcpy(pts + 3 * np++, ps, sizeof(double) * 3);
    while (ri < N) {
      const double* nx = ps + 3 * idx;
      const double df[3] = {nx[0] - cur[0], nx[1] - cur[1], nx[2] - cur[2]};
      const double dn = sqrt((df[0] * df[0] + df[1] * df[1]) + df[2] * df[2]);
      if (dn > limit) {
        for (int a = 0; a < 3; ++a) cur[a] = cur[a] + limit * df[a] / dn;
        memcpy(pts + 3 * np++, cur, sizeof cur);
        ++ri;
        limit = fmax(0.0, samp - P->path_vel_dec * P->dt);
      } else {
        cur[0] = nx[0], cur[1] = nx[1], cur[2] = nx[2];
        if (++idx == ns) {
          for (int i = ri; i < N; ++i) memcpy(pts + 3 * np++, ps + 3 * (ns - 1), sizeof(double) * 3);
          break;
        }
        limit = limit - dn;
      }
    }
  }
  /* KeepOnlyFreeReference (:1665-1693): from the first unknown / occupied sample on, repeat the last free one */
  for (int i = 1; i < np; ++i) {
    const double* pt = pts + 3 * i;
    const int v = vg_get(data, idim, (int)((pt[0] - origin[0]) / P->voxel), (int)((pt[1] - origin[1]) / P->voxel),
                         (int)((pt[2] - origin[2]) / P->voxel));
    if (v == RT_UNKNOWN || v == RT_OCC) {
      for (int j = i; j < np; ++j) memcpy(pts + 3 * j, pts + 3 * (i - 1), sizeof(double) * 3);
      break;
    }
  }
  /* velocity reference (:1528-1546): points backwards along the path, like the reference */
  memset(ref, 0, sizeof(double) * 6 * (size_t)(N + 1));
  double v3[3] = {0, 0, 0};
  for (int i = 0; i < np && i <= N; ++i) memcpy(ref + 6 * i, pts + 3 * i, sizeof(double) * 3);
  if (np > 1) {
    for (int i = 0; i < np - 1 && i <= N; ++i) {
      const double dd = dist3(pts + 3 * i, pts + 3 * (i + 1));
      for (int a = 0; a < 3; ++a) v3[a] = dd > 1e-2 ? vel * (pts[3 * i + a] - pts[3 * (i + 1) + a]) / dd : 0.0;
      memcpy(ref + 6 * i + 3, v3, sizeof v3);
    }
    if (np - 1 <= N) memcpy(ref + 6 * (np - 1) + 3, v3, sizeof v3);
  }
  *path_vel_out = vel;
  free(ps);
  free(pts);
}

typedef struct {
  const reftraj_params* P;
  int n, tid, nt;
  const int8_t* grids;
  size_t grid_stride;
  const int32_t *grid_index, *dims, *n_path, *global_id, *nbr_begin, *nbr_end;
  const double *origins, *path, *prev_ref, *traj, *all_pos;
  const uint8_t *have_prev, *increment, *all_valid;
  int n_rob;
  double *ref, *path_vel;
} rt_job;

static void* rt_worker(void* arg) {
  const rt_job* J = (const rt_job*)arg;
  const reftraj_params* P = J->P;
  const size_t N1 = (size_t)P->n_hor + 1;
  for (int i = J->tid; i < J->n; i += J->nt) {
    const int gi = J->grid_index ? J->grid_index[i] : i;
    rt_generate(P, J->grids + (size_t)gi * J->grid_stride, J->dims + 3 * i, J->origins + 3 * i, J->path + (size_t)i * P->max_path * 3,
                J->n_path[i], J->prev_ref + i * N1 * 3, J->have_prev[i], J->increment[i], J->traj + (size_t)i * P->n_traj * 3,
                J->all_pos, J->all_valid, J->nbr_begin ? J->nbr_begin[i] : 0, J->nbr_end ? J->nbr_end[i] : J->n_rob,
                J->global_id[i], J->ref + i * N1 * 6, J->path_vel + i);
  }
  return 0;
}

int rt_generate_batch(const reftraj_params* P, int n, const int8_t* grids, size_t grid_stride, const int32_t* grid_index,
                      const int32_t* dims, const double* origins, const double* path, const int32_t* n_path,
                      const double* prev_ref, const uint8_t* have_prev, const uint8_t* increment, const double* traj,
                      const int32_t* global_id, const int32_t* nbr_begin, const int32_t* nbr_end, const double* all_pos,
                      const uint8_t* all_valid, int n_rob, double* ref, double* path_vel, int n_threads) {
  if (n_threads < 1) n_threads = 1;
  if (n_threads > 256) n_threads = 256;
  rt_job jobs[256];
  pthread_t th[256];
  for (int t = 0; t < n_threads; ++t) {
    rt_job j = {P, n, t, n_threads, grids, grid_stride, grid_index, dims, n_path, global_id, nbr_begin, nbr_end,
                origins, path, prev_ref, traj, all_pos, have_prev, increment, all_valid, n_rob, ref, path_vel};
    jobs[t] = j;
  }
  for (int t = 1; t < n_threads; ++t) pthread_create(&th[t], 0, rt_worker, &jobs[t]);
  rt_worker(&jobs[0]);
  for (int t = 1; t < n_threads; ++t) pthread_join(th[t], 0);
  return 0;
}
