/* corridor_oracle.c - CPU restatement of the reference's safe-corridor generator.
 *
 * TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
 * leg; the product (multi_agent_pkgs_b200/) never links or calls it.
 *
 * Restates, in plain C:
 *   cor_poly_octa      convex_decomp_lib::GetPolyOcta3D        convex_decomp_util/src/convex_decomp.cpp:5-376
 *   cor_safe_corridor  Agent::GenerateSafeCorridor             multi_agent_planner/src/agent_class.cpp:1236-1447
 *                      (+ LinearConstraint::inside             decomp_geometry/polyhedron.h:130-137,
 *                         VoxelGrid::OccupyUnknown/IsOccupied  voxel_grid_util/src/voxel_grid.cpp:234-240, :150-155)
 *
 * Parity: cor_poly_octa is PINNED - tests/test_corridor_oracle.py compares its planes and its voxel marks
 * bit for bit with the reference's own GetPolyOcta3D (compiled unmodified into oracle/_ref/, see
 * oracle/Makefile target `ref`) live when oracle/_ref exists, and against tests/golden/corridor_*.npz
 * (generated from oracle/_ref by tests/golden/make_corridor_golden.py) everywhere.  cor_safe_corridor
 * (the path walk that picks the seeds) follows agent_class.cpp by reading: that file needs ROS2 and
 * Gurobi headers and cannot be compiled here, so the walk is unpinned; its floating-point expressions
 * are written in the reference's evaluation order (Eigen sums 3-vectors left to right).
 *
 * The shape-aware variant GetPolyOcta3DNew (convex_decomp.cpp:211-...; used by the reference when the seed
 * voxel is squeezed between two occupied voxels, agent_class.cpp:1385-1395) is not restated: such
 * polytopes are generated with the original method and COR_FLAG_SQUEEZED is raised for the agent.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define COR_OCC 100     /* CVX_DCMP_OCC / ENV_BUILDER_OCC */
#define COR_UNKNOWN (-1) /* ENV_BUILDER_UNK */
#define COR_MAX_PLANES 18

#define COR_FLAG_SQUEEZED 1  /* a seed was squeezed: the reference would have switched to GetPolyOcta3DNew */
#define COR_FLAG_ROWS 2      /* a polytope has more rows than rmax (cannot happen for rmax >= 18) */
#define COR_FLAG_SEED_OUT 4  /* a seed fell outside the grid: generation stopped for this agent */
#define COR_FLAG_WINDOW 8    /* (GPU only) the convex set left the 32^3 window around its seed */

typedef struct { int c[3]; } cell_t;

/* outward direction of the six faces (convex_decomp.cpp:15-16) */
static const int kOut[6][3] = {{0, -1, 0}, {1, 0, 0}, {0, 1, 0}, {-1, 0, 0}, {0, 0, 1}, {0, 0, -1}};
/* the two in-face growth axes of each face; directions 2 and 3 are their negatives (:33-38, :63-66) */
static const int kAxes[6][2][3] = {{{1, 0, 0}, {0, 0, 1}},  {{0, 1, 0}, {0, 0, 1}},  {{-1, 0, 0}, {0, 0, 1}},
                                   {{0, -1, 0}, {0, 0, 1}}, {{0, -1, 0}, {1, 0, 0}}, {{0, -1, 0}, {-1, 0, 0}}};
/* box edge met by face f when growing in in-face direction j; the face across that edge; which of that
 * face's four limits points along f's outward direction (:17-25) */
static const int kEdge[6][4] = {{0, 1, 2, 3}, {8, 5, 0, 4}, {10, 9, 8, 11}, {2, 6, 10, 7}, {1, 5, 9, 6}, {3, 7, 11, 4}};
static const int kAcross[6][4] = {{1, 4, 3, 5}, {2, 4, 0, 5}, {3, 4, 1, 5}, {0, 4, 2, 5}, {0, 1, 2, 3}, {0, 3, 2, 1}};
static const int kAcrossLim[6][4] = {{2, 0, 0, 0}, {2, 1, 0, 3}, {2, 2, 0, 2}, {2, 3, 0, 1}, {1, 1, 1, 1}, {3, 3, 3, 3}};
/* the two faces that meet in each of the twelve edges (:26-30) */
static const int kEdgeFaces[12][2] = {{0, 1}, {0, 4}, {0, 3}, {0, 5}, {1, 5}, {1, 4}, {3, 4}, {3, 5}, {1, 2}, {2, 4}, {2, 3}, {2, 5}};

/* chamfer state of one box edge: Corner3D, convex_decomp.hpp:22-36 */
typedef struct {
  double pos[3];
  int slope, dir, fixed, steps;
} edge_t;

typedef struct {
  cell_t* v;
  int n, cap;
} list_t;

static void list_push(list_t* l, cell_t c) {
  if (l->n == l->cap) {
    l->cap = l->cap ? 2 * l->cap : 64;
    l->v = (cell_t*)realloc(l->v, sizeof(cell_t) * (size_t)l->cap);
  }
  l->v[l->n++] = c;
}

/* double-ended list with room on both sides (std::deque in the reference) */
#define DQ_CAP 512
typedef struct {
  cell_t v[DQ_CAP];
  int lo, hi; /* [lo, hi) */
} deq_t;
static void dq_one(deq_t* d, cell_t c) { d->lo = DQ_CAP / 2, d->hi = d->lo + 1, d->v[d->lo] = c; }
static int dq_len(const deq_t* d) { return d->hi - d->lo; }
static void dq_back(deq_t* d, cell_t c) { d->v[d->hi++] = c; }
static void dq_front(deq_t* d, cell_t c) { d->v[--d->lo] = c; }

static int dot3(const int a[3], const int b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static int same(cell_t a, cell_t b) { return a.c[0] == b.c[0] && a.c[1] == b.c[1] && a.c[2] == b.c[2]; }
static cell_t shifted(cell_t a, const int d[3], int s) {
  cell_t r = {{a.c[0] + s * d[0], a.c[1] + s * d[1], a.c[2] + s * d[2]}};
  return r;
}

/* Voxel value with the grid edge acting as occupied.  The reference indexes the array unchecked
 * (:141-150); inside a consistent run every such access is in range, so the guard never changes a result. */
static int vox(const int8_t* data, const int dim[3], cell_t a) {
  if (a.c[0] < 0 || a.c[1] < 0 || a.c[2] < 0 || a.c[0] >= dim[0] || a.c[1] >= dim[1] || a.c[2] >= dim[2]) return COR_OCC;
  return data[a.c[0] + a.c[1] * dim[0] + a.c[2] * dim[0] * dim[1]];
}

/* GetPolyOcta3D: grows an axis-aligned voxel box around `seed`, one face layer at a time in the cyclic
 * order -y, +x, +y, -x, +z, -z, letting each of the twelve box edges degenerate into a chamfer of
 * integer slope.  `data` is modified in place (voxels of the set := conv).  points / normals receive the
 * hyperplanes (chamfers in edge order, then the six faces); returns their number. */
int cor_poly_octa(const int32_t seed_in[3], int8_t* data, const int32_t dim_in[3], int n_it, double res, int conv,
                  const double origin[3], double* points, double* normals) {
  const int dim[3] = {dim_in[0], dim_in[1], dim_in[2]};
  const cell_t seed = {{seed_in[0], seed_in[1], seed_in[2]}};
  list_t cells[6];
  int lim[6][4], alive[6];
  cell_t tip[6];
  edge_t edge[12];
  memset(edge, 0, sizeof edge);
  for (int e = 0; e < 12; ++e) edge[e].dir = -1;
  for (int f = 0; f < 6; ++f) {
    memset(&cells[f], 0, sizeof(list_t));
    list_push(&cells[f], seed);
    tip[f] = seed, alive[f] = 1;
    lim[f][0] = dot3(seed.c, kAxes[f][0]), lim[f][1] = dot3(seed.c, kAxes[f][1]);
    lim[f][2] = -lim[f][0], lim[f][3] = -lim[f][1];  /* :40-43 */
  }
  data[seed.c[0] + seed.c[1] * dim[0] + seed.c[2] * dim[0] * dim[1]] = (int8_t)conv;

  deq_t* ring = (deq_t*)malloc(sizeof(deq_t) * 8); /* ring[0..3]: in-face front lines; top[0..3]: their part above the set */
  deq_t* top = ring + 4;
  list_t layer = {0, 0, 0};
  cell_t* line2 = (cell_t*)malloc(sizeof(cell_t) * DQ_CAP * 2);
  cell_t* liner = line2 + DQ_CAP;

  for (int it = 0; it < n_it; ++it) {
    const int f = it % 6;
    if (!alive[f]) continue;
    const int* out = kOut[f];
    int d4[4][3];
    for (int a = 0; a < 3; ++a) {
      d4[0][a] = kAxes[f][0][a], d4[1][a] = kAxes[f][1][a];
      d4[2][a] = -kAxes[f][0][a], d4[3][a] = -kAxes[f][1][a];
    }
    /* limits of this layer: the face's own, pulled in where a chamfer is running (:70-91) */
    int lm[4];
    edge_t et[4];
    for (int j = 0; j < 4; ++j) {
      lm[j] = lim[f][j];
      et[j] = edge[kEdge[f][j]];
      if (et[j].slope > 0) {
        if (et[j].fixed) {
          if (et[j].dir != f) {
            if (et[j].steps >= et[j].slope) lm[j] -= 1;
          } else {
            lm[j] -= et[j].slope;
          }
        } else if (et[j].dir == f) {
          lm[j] -= et[j].slope;
        }
      }
    }
    /* first cell of the face whose outward neighbour is a free interior voxel within the limits (:97-117) */
    int found = 0;
    cell_t s2 = seed;
    for (int i = 0; i < cells[f].n && !found; ++i) {
      const cell_t t = shifted(cells[f].v[i], out, 1);
      if (t.c[0] >= 1 && t.c[1] >= 1 && t.c[2] >= 1 && t.c[0] < dim[0] - 1 && t.c[1] < dim[1] - 1 && t.c[2] < dim[2] - 1 &&
          vox(data, dim, t) < COR_OCC && dot3(t.c, d4[0]) <= lm[0] && dot3(t.c, d4[1]) <= lm[1] &&
          dot3(t.c, d4[2]) <= lm[2] && dot3(t.c, d4[3]) <= lm[3])
        s2 = t, found = 1;
    }
    if (!found) continue;

    /* in-layer growth from s2: the four front lines advance in turn until none can (:119-209).  A line
     * that failed once can never advance later (its failing cell stays in it, and limits, marks and
     * occupancy do not change during the layer), so closed lines are skipped instead of re-tested. */
    int open[4] = {1, 1, 1, 1};
    cell_t ext[4] = {s2, s2, s2, s2};
    for (int j = 0; j < 4; ++j) dq_one(&ring[j], s2), dq_one(&top[j], s2);
    layer.n = 0;
    list_push(&layer, s2);
    for (int k = 0; open[0] || open[1] || open[2] || open[3]; ++k) {
      const int j = k % 4, jb = (k + 3) % 4, ja = (k + 1) % 4;
      if (!open[j]) continue;
      int n2 = 0, nr = 0, ok = 1;
      for (int i = ring[j].lo; i < ring[j].hi; ++i) {
        const cell_t t = shifted(ring[j].v[i], d4[j], 1);
        if (dot3(t.c, d4[j]) > lm[j]) {
          ok = 0;
          break;
        }
        if (vox(data, dim, shifted(t, out, -1)) == conv) { /* above the set: must be free */
          if (vox(data, dim, t) < COR_OCC) {
            line2[n2++] = t, liner[nr++] = t;
          } else {
            ok = 0;
            break;
          }
        } else {
          line2[n2++] = t;
        }
      }
      if (!ok) {
        open[j] = 0;
        continue;
      }
      ring[j].lo = DQ_CAP / 2 - n2 / 2, ring[j].hi = ring[j].lo + n2;
      memcpy(ring[j].v + ring[j].lo, line2, sizeof(cell_t) * (size_t)n2);
      top[j].lo = DQ_CAP / 2 - nr / 2, top[j].hi = top[j].lo + nr;
      memcpy(top[j].v + top[j].lo, liner, sizeof(cell_t) * (size_t)nr);
      for (int i = 0; i < nr; ++i) list_push(&layer, liner[i]);
      dq_back(&ring[jb], line2[0]);
      dq_front(&ring[ja], line2[n2 - 1]);
      if (nr > 0) {
        if (same(line2[0], liner[0])) dq_back(&top[jb], line2[0]);
        if (same(line2[n2 - 1], liner[nr - 1])) dq_front(&top[ja], line2[n2 - 1]);
      }
      for (int q = 0; q < 4; ++q)
        if (dq_len(&top[q]) > 0) ext[q] = top[q].v[top[q].lo];
    }

    /* chamfer bookkeeping of the four edges around the face (:217-301) */
    int valid = 1;
    for (int j = 0; j < 4 && valid; ++j) {
      edge_t e = et[j];
      if (dq_len(&top[j]) > 0) {
        const cell_t fr = top[j].v[top[j].lo];
        const int dist = lim[f][j] - dot3(fr.c, d4[j]);
        if (e.slope == 0) {
          if (dist > 0) {
            const int* oa = kOut[kAcross[f][j]];
            for (int a = 0; a < 3; ++a) e.pos[a] = fr.c[a] * res - out[a] * res / 2 + oa[a] * res / 2 + res / 2;
            e.slope = dist, e.steps = dist;
            if (dist > 1) e.dir = f;
          }
        } else if (e.fixed) {
          if (e.dir == f || e.dir == -1) {
            if (dist > e.slope) valid = 0;
          } else if (e.steps >= e.slope) {
            if (dist > 1) valid = 0;
            else e.steps = 1;
          } else {
            if (dist != 0) valid = 0;
            else e.steps += 1;
          }
        } else {
          if (e.dir == -1) {
            if (dist == 0) e.dir = kAcross[f][j], e.steps += 1, e.slope += 1;
            else if (dist == 1) e.fixed = 1;
            else valid = 0;
          } else if (e.dir == f) {
            e.slope = dist, e.fixed = 1;
          } else {
            if (dist == 0) e.slope += 1, e.steps += 1;
            else if (dist == 1) e.fixed = 1, e.steps = 1;
            else valid = 0;
          }
        }
      }
      et[j] = e;
    }
    if (!valid) {
      alive[f] = 0;
      continue;
    }
    /* commit the layer (:311-339) */
    cells[f].n = 0;
    for (int i = 0; i < layer.n; ++i) list_push(&cells[f], layer.v[i]);
    for (int j = 0; j < 4; ++j) {
      lim[f][j] = dot3(ext[j].c, d4[j]);
      edge[kEdge[f][j]] = et[j];
      if (et[j].slope == 0 && dq_len(&top[j]) > 0 && lim[f][j] - dot3(top[j].v[top[j].lo].c, d4[j]) == 0) {
        const int g = kAcross[f][j]; /* the neighbouring face gains this line of cells and one unit of limit */
        for (int i = top[j].lo; i < top[j].hi; ++i) list_push(&cells[g], top[j].v[i]);
        lim[g][kAcrossLim[f][j]] += 1;
      }
    }
    tip[f] = layer.v[0];
    for (int i = 0; i < layer.n; ++i)
      data[layer.v[i].c[0] + layer.v[i].c[1] * dim[0] + layer.v[i].c[2] * dim[0] * dim[1]] = (int8_t)conv;
  }

  /* hyperplanes (:343-375): chamfers, then faces */
  int np = 0;
  for (int e = 0; e < 12; ++e) {
    if (edge[e].slope <= 0) continue;
    const int f1 = kEdgeFaces[e][0], f2 = kEdgeFaces[e][1];
    const int* steep = edge[e].dir == f1 ? kOut[f1] : kOut[f2];
    const int* flat = edge[e].dir == f1 ? kOut[f2] : kOut[f1];
    for (int a = 0; a < 3; ++a) {
      normals[3 * np + a] = (double)(edge[e].slope * steep[a] + flat[a]);
      points[3 * np + a] = edge[e].pos[a] + origin[a];
    }
    ++np;
  }
  for (int f = 0; f < 6; ++f) {
    for (int a = 0; a < 3; ++a) {
      const double p = tip[f].c[a] * res + kOut[f][a] * res / 2 + res / 2;
      points[3 * np + a] = p + origin[a];
      normals[3 * np + a] = (double)kOut[f][a];
    }
    ++np;
  }
  for (int f = 0; f < 6; ++f) free(cells[f].v);
  free(layer.v);
  free(ring);
  free(line2);
  return np;
}

/* ---------------------------------------------------------------------------------------------- */
typedef struct cor_params {
  int32_t poly_hor;   /* poly_hor_ */
  int32_t n_it;       /* n_it_decomp_ */
  int32_t rmax;       /* row stride of the polytope arrays */
  int32_t n_traj;     /* points of the previous plan traj_curr_ (N + 1) */
  int32_t max_path;   /* row stride of the path array */
  int32_t reserved;
  double voxel;       /* voxel size */
} cor_params;

static int inside_rows(const double* A, const double* b, int rows, const double pt[3]) {
  for (int r = 0; r < rows; ++r) /* LinearConstraint::inside: any A x - b > 0 is outside (polyhedron.h:130-137) */
    if (A[3 * r] * pt[0] + A[3 * r + 1] * pt[1] + A[3 * r + 2] * pt[2] - b[r] > 0) return 0;
  return 1;
}

/* Agent::GenerateSafeCorridor for one agent.
 *   grid [dz][dy][dx] int8 (x fastest), dim = (dx, dy, dz), origin: the agent's local voxel grid
 *   pos: state_curr_ position; path [n_path][3]: path_curr_ (global path, current position not included)
 *   prev_*: poly_const_vec_ / poly_seeds_ / poly_used_idx_ of the previous step (prev_n = 0: none);
 *           prev_traj [n_traj][3]: positions of traj_curr_
 *   out: poly_A [P][rmax][3], poly_b [P][rmax], poly_rows [P] (0 = absent), seeds [P][3]; returns flags */
int cor_safe_corridor(const cor_params* P, const int8_t* grid, const int32_t dim[3], const double origin[3],
                      const double pos[3], const double* path, int n_path, int prev_n, const double* prev_A,
                      const double* prev_b, const int32_t* prev_rows, const double* prev_seeds, const uint8_t* prev_used,
                      const double* prev_traj, double* poly_A, double* poly_b, int32_t* poly_rows, double* seeds) {
  const int PH = P->poly_hor, R = P->rmax;
  int n_poly = 0, flags = 0;
  memset(poly_A, 0, sizeof(double) * (size_t)PH * R * 3);
  memset(poly_b, 0, sizeof(double) * (size_t)PH * R);
  memset(poly_rows, 0, sizeof(int32_t) * (size_t)PH);
  memset(seeds, 0, sizeof(double) * (size_t)PH * 3);
#define KEEP(i)                                                                   \
  do {                                                                            \
    memcpy(poly_A + (size_t)n_poly * R * 3, prev_A + (size_t)(i) * R * 3, sizeof(double) * (size_t)R * 3); \
    memcpy(poly_b + (size_t)n_poly * R, prev_b + (size_t)(i) * R, sizeof(double) * (size_t)R);             \
    poly_rows[n_poly] = prev_rows[i];                                             \
    memcpy(seeds + 3 * n_poly, prev_seeds + 3 * (i), sizeof(double) * 3);         \
    ++n_poly;                                                                     \
  } while (0)
  if (prev_n > 0) {
    /* the whole previous plan inside the last polytope: keep only that one (:1252-1266) */
    const int last = prev_n - 1;
    int all_in = 1;
    for (int j = 0; j < P->n_traj && all_in; ++j)
      all_in = inside_rows(prev_A + (size_t)last * R * 3, prev_b + (size_t)last * R, prev_rows[last], prev_traj + 3 * j);
    if (all_in) {
      KEEP(last);
    } else { /* otherwise the polytopes the last optimisation used (:1272-1281) */
      for (int i = 0; i < prev_n; ++i)
        if (prev_used[i]) KEEP(i);
    }
  }
#undef KEEP
  if (n_path < 1) return flags; /* the reference would index past the end of the path here */

  /* OccupyUnknown on a private copy of the grid (:1289-1305) */
  const size_t nvox = (size_t)dim[0] * dim[1] * dim[2];
  int8_t* data = (int8_t*)malloc(2 * nvox);
  int8_t* work = data + nvox;
  for (size_t i = 0; i < nvox; ++i) data[i] = grid[i] == COR_UNKNOWN ? COR_OCC : grid[i];
  const double vs = P->voxel, samp = vs / 10;
  double cur[3] = {pos[0], pos[1], pos[2]};
  int path_idx = 1; /* index into [pos, path...] */
  const int path_len = n_path + 1;
  while (n_poly < PH) {
    const double* next = path + 3 * (path_idx - 1);
    const double diff[3] = {next[0] - cur[0], next[1] - cur[1], next[2] - cur[2]};
    const double dist = sqrt(diff[0] * diff[0] + diff[1] * diff[1] + diff[2] * diff[2]);
    if (dist > samp) {
      for (int a = 0; a < 3; ++a) cur[a] = cur[a] + samp * diff[a] / dist;
    } else {
      for (int a = 0; a < 3; ++a) cur[a] = next[a];
      if (++path_idx == path_len) break;
    }
    int in_any = 0;
    for (int i = 0; i < n_poly && !in_any; ++i)
      in_any = inside_rows(poly_A + (size_t)i * R * 3, poly_b + (size_t)i * R, poly_rows[i], cur);
    if (in_any) continue;
    double sp[3] = {cur[0], cur[1], cur[2]};
    if (dist > 0) {
      const double m = samp < dist ? samp : dist; /* one sample back: the last point still inside (:1343-1346) */
      for (int a = 0; a < 3; ++a) sp[a] = cur[a] - m * diff[a] / dist;
    }
    int32_t sv[3];
    double sw[3];
    for (int a = 0; a < 3; ++a) {
      sv[a] = (int32_t)((sp[a] - origin[a]) / vs);
      sw[a] = sv[a] * vs + vs / 2 + origin[a];
    }
    int seen = 0;
    for (int i = 0; i < n_poly && !seen; ++i) seen = sw[0] == seeds[3 * i] && sw[1] == seeds[3 * i + 1] && sw[2] == seeds[3 * i + 2];
    if (seen) continue;
    if (sv[0] < 0 || sv[1] < 0 || sv[2] < 0 || sv[0] >= dim[0] || sv[1] >= dim[1] || sv[2] >= dim[2]) {
      flags |= COR_FLAG_SEED_OUT;
      break;
    }
    /* squeezed seed: the reference switches to GetPolyOcta3DNew (:1385-1395); IsOccupied is false outside */
    for (int a = 0; a < 3; ++a) {
      cell_t lo = {{sv[0], sv[1], sv[2]}}, hi = lo;
      lo.c[a] -= 1, hi.c[a] += 1;
      const int in_lo = lo.c[a] >= 0, in_hi = hi.c[a] < dim[a];
      if (in_lo && in_hi && vox(data, dim, lo) == COR_OCC && vox(data, dim, hi) == COR_OCC) flags |= COR_FLAG_SQUEEZED;
    }
    double pts[3 * COR_MAX_PLANES], nrm[3 * COR_MAX_PLANES];
    /* a fresh copy of the grid per polytope, like the reference (:1405): the seed voxel is marked even
     * when it is occupied, so marks left behind would read as free space to the next polytope */
    memcpy(work, data, nvox);
    const int np = cor_poly_octa(sv, work, dim, P->n_it, vs, -(n_poly + 1), origin, pts, nrm);
    if (np > R) {
      flags |= COR_FLAG_ROWS;
      break;
    }
    double* A = poly_A + (size_t)n_poly * R * 3;
    double* b = poly_b + (size_t)n_poly * R;
    for (int i = 0; i < np; ++i) { /* A = normal, b = point . normal (:1428-1437) */
      for (int a = 0; a < 3; ++a) A[3 * i + a] = nrm[3 * i + a];
      b[i] = pts[3 * i] * nrm[3 * i] + pts[3 * i + 1] * nrm[3 * i + 1] + pts[3 * i + 2] * nrm[3 * i + 2];
    }
    poly_rows[n_poly] = np;
    for (int a = 0; a < 3; ++a) seeds[3 * n_poly + a] = sw[a];
    ++n_poly;
  }
  free(data);
  return flags;
}

/* ---------------------------------------------------------------------------------------------- */
typedef struct {
  const cor_params* P;
  int n, tid, nt;
  const int8_t* grids;
  const int32_t* dims;
  const double *origins, *pos, *path;
  const int32_t* n_path;
  const int32_t* prev_n;
  const double *prev_A, *prev_b;
  const int32_t* prev_rows;
  const double* prev_seeds;
  const uint8_t* prev_used;
  const double* prev_traj;
  double *poly_A, *poly_b;
  int32_t* poly_rows;
  double* seeds;
  int32_t* flags;
  size_t grid_stride;
} job_t;

static void* worker(void* arg) {
  const job_t* J = (const job_t*)arg;
  const cor_params* P = J->P;
  const size_t PH = (size_t)P->poly_hor, R = (size_t)P->rmax;
  for (int i = J->tid; i < J->n; i += J->nt) {
    J->flags[i] = cor_safe_corridor(
        P, J->grids + (size_t)i * J->grid_stride, J->dims + 3 * i, J->origins + 3 * i, J->pos + 3 * i,
        J->path + (size_t)i * P->max_path * 3, J->n_path[i], J->prev_n ? J->prev_n[i] : 0,
        J->prev_A ? J->prev_A + i * PH * R * 3 : 0, J->prev_b ? J->prev_b + i * PH * R : 0,
        J->prev_rows ? J->prev_rows + i * PH : 0, J->prev_seeds ? J->prev_seeds + i * PH * 3 : 0,
        J->prev_used ? J->prev_used + i * PH : 0, J->prev_traj ? J->prev_traj + (size_t)i * P->n_traj * 3 : 0,
        J->poly_A + i * PH * R * 3, J->poly_b + i * PH * R, J->poly_rows + i * PH, J->seeds + i * PH * 3);
  }
  return 0;
}

/* batch over agents; grids [n][grid_stride] with per-agent dims / origins; arrays as in cor_safe_corridor */
int cor_safe_corridor_batch(const cor_params* P, int n, const int8_t* grids, size_t grid_stride, const int32_t* dims,
                            const double* origins, const double* pos, const double* path, const int32_t* n_path,
                            const int32_t* prev_n, const double* prev_A, const double* prev_b, const int32_t* prev_rows,
                            const double* prev_seeds, const uint8_t* prev_used, const double* prev_traj, double* poly_A,
                            double* poly_b, int32_t* poly_rows, double* seeds, int32_t* flags, int n_threads) {
  if (n_threads < 1) n_threads = 1;
  if (n_threads > 256) n_threads = 256;
  job_t jobs[256];
  pthread_t th[256];
  for (int t = 0; t < n_threads; ++t) {
    job_t j = {P, n, t, n_threads, grids, dims, origins, pos, path, n_path, prev_n, prev_A, prev_b, prev_rows,
               prev_seeds, prev_used, prev_traj, poly_A, poly_b, poly_rows, seeds, flags, grid_stride};
    jobs[t] = j;
  }
  for (int t = 1; t < n_threads; ++t) pthread_create(&th[t], 0, worker, &jobs[t]);
  worker(&jobs[0]);
  for (int t = 1; t < n_threads; ++t) pthread_join(th[t], 0);
  return 0;
}
