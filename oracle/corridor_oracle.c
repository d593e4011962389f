/* corridor_oracle.c - CPU restatement of the reference's safe-corridor generator.
 *
 * TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
 * leg; the product (multi_agent_pkgs_b200/) never links or calls it.
 *
 * Restates, in plain C:
 *   cor_poly_octa      convex_decomp_lib::GetPolyOcta3D        convex_decomp_util/src/convex_decomp.cpp:5-376
 *   cor_poly_octa_new  convex_decomp_lib::GetPolyOcta3DNew     convex_decomp_util/src/convex_decomp.cpp:590-1162
 *   cor_safe_corridor  Agent::GenerateSafeCorridor             multi_agent_planner/src/agent_class.cpp:1236-1447
 *                      (+ LinearConstraint::inside             decomp_geometry/polyhedron.h:130-137,
 *                         VoxelGrid::OccupyUnknown/IsOccupied  voxel_grid_util/src/voxel_grid.cpp:234-240, :150-155)
 *
 * Parity: cor_poly_octa is PINNED - tests/test_corridor_oracle.py compares its planes and its voxel marks
 * bit for bit with the reference's own GetPolyOcta3D (compiled unmodified into oracle/_ref/, see
 * oracle/Makefile target `ref`) live when oracle/_ref exists, and against tests/golden/corridor_*.npz
 * (generated from oracle/_ref by tests/golden/make_corridor_golden.py) everywhere.  cor_safe_corridor
 * (the path walk that picks the seeds) is PINNED as well: agent_class.cpp compiles unmodified on the stand-in
 * ROS / Gurobi headers of oracle/ref_shim/ (oracle/_ref/libref_agent.so) and its own GenerateSafeCorridor gives
 * bit-identical polytopes and seeds, first and follow-up updates (tests/test_ref_agent.py, fixture
 * tests/golden/corridor_node_ref.npz).  Floating-point expressions are written in the reference's evaluation
 * order (Eigen sums 3-vectors left to right).
 *
 * The shape-aware variant GetPolyOcta3DNew (convex_decomp.cpp:590-1162 with FindCorners :378-561 and
 * SideIsEmpty :197-209; used by the reference when the seed voxel is squeezed between two occupied voxels,
 * agent_class.cpp:1385-1395, or when use_cvx_new is set) is cor_poly_octa_new, pinned the same way
 * (3 600 random grids, a third of them with squeezed seeds).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define COR_OCC 100     /* CVX_DCMP_OCC / ENV_BUILDER_OCC */
#define COR_UNKNOWN (-1) /* ENV_BUILDER_UNK */
#define COR_MAX_PLANES 18

#define COR_FLAG_SQUEEZED 1  /* informational: a seed was squeezed and GetPolyOcta3DNew was used, like the reference */
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

/* ------------------------------------------------------------------------------------------------
 * Shared machinery of GetPolyOcta3D (convex_decomp.cpp:5-376), GetPolyOcta3DNew (:590-1162) and
 * FindCorners (:378-561): the state of the growing box, the limits of a layer, the search for the
 * layer's first cell and the in-layer growth of the four front lines.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  list_t cells[6];
  int lim[6][4], alive[6];
  cell_t tip[6];
  edge_t edge[12];
} box_t;

typedef struct {
  deq_t ring[4], top[4]; /* in-face front lines; their part above the set (borders_2d / borders_2d_real) */
  list_t layer;          /* border_real_tmp */
  cell_t ext[4];         /* border_limit_tmp: front of top[q] the last time it was non-empty */
  int lm[4];             /* borders_limits */
  edge_t et[4];          /* corners_tmp */
  int d4[4][3];
} layer_t;

static void face_dirs(int f, int d4[4][3]) {
  for (int a = 0; a < 3; ++a) {
    d4[0][a] = kAxes[f][0][a], d4[1][a] = kAxes[f][1][a];
    d4[2][a] = -kAxes[f][0][a], d4[3][a] = -kAxes[f][1][a];
  }
}

/* limits of the next layer of face f: the face's own, pulled in where a chamfer is running (:70-91) */
static void layer_limits(const int lim[4], const edge_t* edge, int f, layer_t* L) {
  for (int j = 0; j < 4; ++j) {
    L->lm[j] = lim[j];
    L->et[j] = edge[kEdge[f][j]];
    const edge_t* e = &L->et[j];
    if (e->slope > 0) {
      if (e->fixed) {
        if (e->dir != f) {
          if (e->steps >= e->slope) L->lm[j] -= 1;
        } else {
          L->lm[j] -= e->slope;
        }
      } else if (e->dir == f) {
        L->lm[j] -= e->slope;
      }
    }
  }
}

/* first cell of the face whose outward neighbour is a free voxel inside [1, dim - margin) within the limits
 * (:97-117; margin = 1 in GetPolyOcta3D, 0 in GetPolyOcta3DNew / FindCorners) */
static int find_seed2d(const list_t* cells, int f, const layer_t* L, const int8_t* data, const int dim[3], int margin, cell_t* s2) {
  for (int i = 0; i < cells->n; ++i) {
    const cell_t t = shifted(cells->v[i], kOut[f], 1);
    if (t.c[0] >= 1 && t.c[1] >= 1 && t.c[2] >= 1 && t.c[0] < dim[0] - margin && t.c[1] < dim[1] - margin &&
        t.c[2] < dim[2] - margin && vox(data, dim, t) < COR_OCC && dot3(t.c, L->d4[0]) <= L->lm[0] &&
        dot3(t.c, L->d4[1]) <= L->lm[1] && dot3(t.c, L->d4[2]) <= L->lm[2] && dot3(t.c, L->d4[3]) <= L->lm[3]) {
      *s2 = t;
      return 1;
    }
  }
  return 0;
}

/* In-layer growth from s2: the four front lines advance in turn until none can (:119-209).  A line that
 * failed once can never advance later (its failing cell stays in it, and limits, marks and occupancy do
 * not change during the layer), so closed lines are skipped instead of re-tested. */
static void grow_lines(int f, cell_t s2, layer_t* L, const int8_t* data, const int dim[3], int conv, cell_t* line2, cell_t* liner) {
  const int* out = kOut[f];
  int open[4] = {1, 1, 1, 1};
  for (int j = 0; j < 4; ++j) dq_one(&L->ring[j], s2), dq_one(&L->top[j], s2), L->ext[j] = s2;
  L->layer.n = 0;
  list_push(&L->layer, s2);
  for (int k = 0; open[0] || open[1] || open[2] || open[3]; ++k) {
    const int j = k % 4, jb = (k + 3) % 4, ja = (k + 1) % 4;
    if (!open[j]) continue;
    int n2 = 0, nr = 0, ok = 1;
    for (int i = L->ring[j].lo; i < L->ring[j].hi; ++i) {
      const cell_t t = shifted(L->ring[j].v[i], L->d4[j], 1);
      if (dot3(t.c, L->d4[j]) > L->lm[j]) {
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
    L->ring[j].lo = DQ_CAP / 2 - n2 / 2, L->ring[j].hi = L->ring[j].lo + n2;
    memcpy(L->ring[j].v + L->ring[j].lo, line2, sizeof(cell_t) * (size_t)n2);
    L->top[j].lo = DQ_CAP / 2 - nr / 2, L->top[j].hi = L->top[j].lo + nr;
    memcpy(L->top[j].v + L->top[j].lo, liner, sizeof(cell_t) * (size_t)nr);
    for (int i = 0; i < nr; ++i) list_push(&L->layer, liner[i]);
    dq_back(&L->ring[jb], line2[0]);
    dq_front(&L->ring[ja], line2[n2 - 1]);
    if (nr > 0) {
      if (same(line2[0], liner[0])) dq_back(&L->top[jb], line2[0]);
      if (same(line2[n2 - 1], liner[nr - 1])) dq_front(&L->top[ja], line2[n2 - 1]);
    }
    for (int q = 0; q < 4; ++q)
      if (dq_len(&L->top[q]) > 0) L->ext[q] = L->top[q].v[L->top[q].lo];
  }
}

static void box_init(box_t* B, cell_t seed) {
  memset(B->edge, 0, sizeof B->edge);
  for (int e = 0; e < 12; ++e) B->edge[e].dir = -1;
  for (int f = 0; f < 6; ++f) {
    memset(&B->cells[f], 0, sizeof(list_t));
    list_push(&B->cells[f], seed);
    B->tip[f] = seed, B->alive[f] = 1;
    B->lim[f][0] = dot3(seed.c, kAxes[f][0]), B->lim[f][1] = dot3(seed.c, kAxes[f][1]);
    B->lim[f][2] = -B->lim[f][0], B->lim[f][3] = -B->lim[f][1]; /* :40-43 */
  }
}

/* commit a layer: the face takes it over, its limits move, neighbouring faces gain the lines that touch
 * them (:311-339) */
static void commit_layer(box_t* B, int f, const layer_t* L, int8_t* data, const int dim[3], int conv) {
  B->cells[f].n = 0;
  for (int i = 0; i < L->layer.n; ++i) list_push(&B->cells[f], L->layer.v[i]);
  for (int j = 0; j < 4; ++j) {
    B->lim[f][j] = dot3(L->ext[j].c, L->d4[j]);
    B->edge[kEdge[f][j]] = L->et[j];
    if (L->et[j].slope == 0 && dq_len(&L->top[j]) > 0 && B->lim[f][j] - dot3(L->top[j].v[L->top[j].lo].c, L->d4[j]) == 0) {
      const int g = kAcross[f][j];
      for (int i = L->top[j].lo; i < L->top[j].hi; ++i) list_push(&B->cells[g], L->top[j].v[i]);
      B->lim[g][kAcrossLim[f][j]] += 1;
    }
  }
  B->tip[f] = L->layer.v[0];
  for (int i = 0; i < L->layer.n; ++i)
    data[L->layer.v[i].c[0] + L->layer.v[i].c[1] * dim[0] + L->layer.v[i].c[2] * dim[0] * dim[1]] = (int8_t)conv;
}

/* hyperplanes (:343-375): chamfers in edge order, then the six faces */
static int box_planes(const box_t* B, double res, const double origin[3], double* points, double* normals) {
  int np = 0;
  for (int e = 0; e < 12; ++e) {
    if (B->edge[e].slope <= 0) continue;
    const int f1 = kEdgeFaces[e][0], f2 = kEdgeFaces[e][1];
    const int* steep = B->edge[e].dir == f1 ? kOut[f1] : kOut[f2];
    const int* flat = B->edge[e].dir == f1 ? kOut[f2] : kOut[f1];
    for (int a = 0; a < 3; ++a) {
      normals[3 * np + a] = (double)(B->edge[e].slope * steep[a] + flat[a]);
      points[3 * np + a] = B->edge[e].pos[a] + origin[a];
    }
    ++np;
  }
  for (int f = 0; f < 6; ++f) {
    for (int a = 0; a < 3; ++a) {
      const double p = B->tip[f].c[a] * res + kOut[f][a] * res / 2 + res / 2;
      points[3 * np + a] = p + origin[a];
      normals[3 * np + a] = (double)kOut[f][a];
    }
    ++np;
  }
  return np;
}

static layer_t* layer_new(void) {
  layer_t* L = (layer_t*)calloc(1, sizeof(layer_t));
  return L;
}
static void layer_free(layer_t* L) {
  free(L->layer.v);
  free(L);
}

/* GetPolyOcta3D: grows an axis-aligned voxel box around `seed`, one face layer at a time in the cyclic
 * order -y, +x, +y, -x, +z, -z, letting each of the twelve box edges degenerate into a chamfer of
 * integer slope.  `data` is modified in place (voxels of the set := conv).  points / normals receive the
 * hyperplanes (chamfers in edge order, then the six faces); returns their number. */
int cor_poly_octa(const int32_t seed_in[3], int8_t* data, const int32_t dim_in[3], int n_it, double res, int conv,
                  const double origin[3], double* points, double* normals) {
  const int dim[3] = {dim_in[0], dim_in[1], dim_in[2]};
  const cell_t seed = {{seed_in[0], seed_in[1], seed_in[2]}};
  box_t B;
  box_init(&B, seed);
  data[seed.c[0] + seed.c[1] * dim[0] + seed.c[2] * dim[0] * dim[1]] = (int8_t)conv;
  layer_t* L = layer_new();
  cell_t* line2 = (cell_t*)malloc(sizeof(cell_t) * DQ_CAP * 2);
  cell_t* liner = line2 + DQ_CAP;

  for (int it = 0; it < n_it; ++it) {
    const int f = it % 6;
    if (!B.alive[f]) continue;
    const int* out = kOut[f];
    face_dirs(f, L->d4);
    layer_limits(B.lim[f], B.edge, f, L);
    cell_t s2;
    if (!find_seed2d(&B.cells[f], f, L, data, dim, 1, &s2)) continue;
    grow_lines(f, s2, L, data, dim, conv, line2, liner);

    /* chamfer bookkeeping of the four edges around the face (:217-301) */
    int valid = 1;
    for (int j = 0; j < 4 && valid; ++j) {
      edge_t e = L->et[j];
      if (dq_len(&L->top[j]) > 0) {
        const cell_t fr = L->top[j].v[L->top[j].lo];
        const int dist = B.lim[f][j] - dot3(fr.c, L->d4[j]);
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
      L->et[j] = e;
    }
    if (!valid) {
      B.alive[f] = 0;
      continue;
    }
    commit_layer(&B, f, L, data, dim, conv);
  }
  const int np = box_planes(&B, res, origin, points, normals);
  for (int f = 0; f < 6; ++f) free(B.cells[f].v);
  layer_free(L);
  free(line2);
  return np;
}

/* ------------------------------------------------------------------------------------------------
 * GetPolyOcta3DNew (:590-1162), the shape-aware variant the reference switches to when the seed voxel is
 * squeezed between two occupied voxels (agent_class.cpp:1385-1395).  Differences from GetPolyOcta3D: the
 * layer's first cell may lie in the last voxel plane; a layer that would shrink the face to less than half
 * its area is refused without closing the face; when an edge starts a chamfer, the layer is accepted only
 * if there is something to chamfer around (SideIsEmpty, :197-209) and a trial of the following layer
 * (FindCorners, :378-561) does not contradict the slope just chosen.
 * ---------------------------------------------------------------------------------------------- */
static double line_area(const layer_t* L) {
  /* | |front0 . d0| - |front2 . d2| | * | |front1 . d1| - |front3 . d3| |  (:214-218, :518-522).  The reference
   * calls front() on lines that may be empty; with libstdc++ that reads the stale first slot, which is the
   * front the line had the last time it was non-empty - exactly border_limit_tmp (ext). */
  const double a0 = fabs((double)dot3(L->ext[0].c, L->d4[0])), a2 = fabs((double)dot3(L->ext[2].c, L->d4[2]));
  const double a1 = fabs((double)dot3(L->ext[1].c, L->d4[1])), a3 = fabs((double)dot3(L->ext[3].c, L->d4[3]));
  return fabs(a0 - a2) * fabs(a1 - a3);
}

static int get_voxel(const int8_t* data, const int dim[3], cell_t a) { return vox(data, dim, a); } /* GetVoxel :187-195 */

/* SideIsEmpty (:197-209): a non-empty list of cells whose neighbours in direction inc all hold a value <= 0 */
static int side_is_empty(const cell_t* v, int n, const int inc[3], const int8_t* data, const int dim[3]) {
  if (n == 0) return 0;
  for (int i = 0; i < n; ++i)
    if (get_voxel(data, dim, shifted(v[i], inc, 1)) > 0) return 0;
  return 1;
}

/* FindCorners (:378-561): trial of the next layer of face f on the boxes `cells` / `lim`; fills et_out with the
 * edge states it would produce (only "a chamfer starts" is evaluated) and clears *valid when the face is
 * closed or the trial layer would shrink to less than half the area. */
static void find_corners(int f, const int alive[6], const list_t cells[6], const int lim[6][4], const edge_t* edge,
                         const int8_t* data, const int dim[3], int conv, layer_t* T, cell_t* line2, cell_t* liner,
                         edge_t et_out[4], int* valid) {
  if (!alive[f]) {
    *valid = 0;
    return;
  }
  face_dirs(f, T->d4);
  layer_limits(lim[f], edge, f, T);
  for (int j = 0; j < 4; ++j) et_out[j] = T->et[j];
  const double area = fabs(fabs((double)T->lm[0]) - fabs((double)T->lm[2])) * fabs(fabs((double)T->lm[1]) - fabs((double)T->lm[3]));
  cell_t s2;
  if (!find_seed2d(&cells[f], f, T, data, dim, 0, &s2)) return;
  grow_lines(f, s2, T, data, dim, conv, line2, liner);
  if (line_area(T) < area / 2) *valid = 0;
  for (int j = 0; j < 4; ++j) {
    edge_t e = et_out[j];
    if (dq_len(&T->top[j]) > 0) {
      const int dist = lim[f][j] - dot3(T->top[j].v[T->top[j].lo].c, T->d4[j]);
      if (e.slope == 0 && dist > 0) {
        e.slope = dist, e.steps = dist;
        if (dist > 1) e.dir = f;
      }
    }
    et_out[j] = e;
  }
}

int cor_poly_octa_new(const int32_t seed_in[3], int8_t* data, const int32_t dim_in[3], int n_it, double res, int conv,
                      const double origin[3], double* points, double* normals) {
  const int dim[3] = {dim_in[0], dim_in[1], dim_in[2]};
  const cell_t seed = {{seed_in[0], seed_in[1], seed_in[2]}};
  box_t B;
  box_init(&B, seed);
  data[seed.c[0] + seed.c[1] * dim[0] + seed.c[2] * dim[0] * dim[1]] = (int8_t)conv;
  layer_t* L = layer_new();
  layer_t* T = layer_new();
  cell_t* line2 = (cell_t*)malloc(sizeof(cell_t) * DQ_CAP * 2);
  cell_t* liner = line2 + DQ_CAP;
  list_t side = {0, 0, 0}, tmp_cells = {0, 0, 0};
  int8_t* saved = 0;
  int saved_cap = 0;

  for (int it = 0; it < n_it; ++it) {
    const int f = it % 6;
    if (!B.alive[f]) continue;
    const int* out = kOut[f];
    face_dirs(f, L->d4);
    layer_limits(B.lim[f], B.edge, f, L);
    const double area = fabs(fabs((double)L->lm[0]) - fabs((double)L->lm[2])) * fabs(fabs((double)L->lm[1]) - fabs((double)L->lm[3]));
    cell_t s2;
    if (!find_seed2d(&B.cells[f], f, L, data, dim, 0, &s2)) continue;
    grow_lines(f, s2, L, data, dim, conv, line2, liner);
    int soft = !(line_area(L) < area / 2);

    /* chamfer bookkeeping (:228-327) */
    int valid = 1, started[4] = {0, 0, 0, 0};
    for (int j = 0; j < 4; ++j) {
      edge_t e = L->et[j];
      int stop = 0;
      if (dq_len(&L->top[j]) > 0) {
        const cell_t fr = L->top[j].v[L->top[j].lo];
        const int dist = B.lim[f][j] - dot3(fr.c, L->d4[j]);
        if (e.slope == 0) {
          if (dist > 0) {
            const int* oa = kOut[kAcross[f][j]];
            for (int a = 0; a < 3; ++a) e.pos[a] = fr.c[a] * res - out[a] * res / 2 + oa[a] * res / 2 + res / 2 + res / 2;
            e.slope = dist, e.steps = dist;
            if (dist > 1) e.dir = f, started[j] = 2; /* check one side */
            else started[j] = 1;                      /* check both sides */
            if (it < 6) soft = 0;                     /* no chamfers during the first round */
          }
        } else if (e.fixed) {
          if (e.dir == f || e.dir == -1) {
            if (dist > e.slope) stop = 1; /* the reference leaves the loop here WITHOUT invalidating the layer (:269-271) */
          } else if (e.steps >= e.slope) {
            if (dist > 1) valid = 0, stop = 1;
            else e.steps = 1;
          } else {
            if (dist != 0) valid = 0, stop = 1;
            else e.steps += 1;
          }
        } else {
          if (e.dir == -1) {
            if (dist == 0) e.dir = kAcross[f][j], e.steps += 1, e.slope += 1;
            else if (dist == 1) e.fixed = 1;
            else valid = 0, stop = 1;
          } else if (e.dir == f) {
            e.slope = dist, e.fixed = 1;
          } else {
            if (dist == 0) e.slope += 1, e.steps += 1;
            else if (dist == 1) e.fixed = 1, e.steps = 1;
            else valid = 0, stop = 1;
          }
        }
      }
      if (stop) break; /* corners_tmp[j..3] keep their copies */
      L->et[j] = e;
    }
    if (!(valid && soft)) {
      if (!valid) B.alive[f] = 0;
      continue;
    }

    /* is there anything to chamfer around?  (:330-366) */
    int expand = 1;
    for (int j = 0; j < 4; ++j) {
      if (!started[j]) continue;
      const int first = side_is_empty(L->top[j].v + L->top[j].lo, dq_len(&L->top[j]), out, data, dim);
      const int g = kAcross[f][j], q = kAcrossLim[f][j];
      int dg[4][3];
      face_dirs(g, dg);
      side.n = 0;
      for (int i = 0; i < B.cells[g].n; ++i)
        if (dot3(B.cells[g].v[i].c, dg[q]) == B.lim[g][q]) list_push(&side, B.cells[g].v[i]);
      int second = 1;
      if (started[j] == 1) second = side_is_empty(side.v, side.n, kOut[g], data, dim);
      expand = !(first && second);
      if (!expand) {
        B.alive[f] = 0;
        break;
      }
    }
    if (expand && (started[0] || started[1] || started[2] || started[3])) {
      /* trial: mark the layer, look one layer further on this face, unmark (:369-413) */
      if (L->layer.n > saved_cap) saved_cap = 2 * L->layer.n, saved = (int8_t*)realloc(saved, (size_t)saved_cap);
      for (int i = 0; i < L->layer.n; ++i) {
        int8_t* d = &data[L->layer.v[i].c[0] + L->layer.v[i].c[1] * dim[0] + L->layer.v[i].c[2] * dim[0] * dim[1]];
        saved[i] = *d, *d = (int8_t)conv;
      }
      list_t cells_t[6];
      int lim_t[6][4];
      memcpy(cells_t, B.cells, sizeof cells_t);
      memcpy(lim_t, B.lim, sizeof lim_t);
      tmp_cells.n = 0;
      for (int i = 0; i < L->layer.n; ++i) list_push(&tmp_cells, L->layer.v[i]);
      cells_t[f] = tmp_cells;
      for (int j = 0; j < 4; ++j) lim_t[f][j] = dot3(L->ext[j].c, L->d4[j]);
      edge_t edge_t_[12];
      memcpy(edge_t_, B.edge, sizeof edge_t_);
      for (int j = 0; j < 4; ++j)
        if (started[j] == 0) edge_t_[kEdge[f][j]] = L->et[j];
      int vfinal = 1;
      edge_t fin[4];
      find_corners(f, B.alive, cells_t, lim_t, edge_t_, data, dim, conv, T, line2, liner, fin, &vfinal);
      for (int i = 0; i < L->layer.n; ++i)
        data[L->layer.v[i].c[0] + L->layer.v[i].c[1] * dim[0] + L->layer.v[i].c[2] * dim[0] * dim[1]] = saved[i];
      if (vfinal) {
        for (int j = 0; j < 4; ++j)
          if (started[j] == 2 && fin[j].slope < L->et[j].slope) {
            expand = 0, B.alive[f] = 0;
            break;
          }
        if (expand) {
          for (int j = 0; j < 4; ++j) {
            if (started[j] != 1) continue;
            const int g = kAcross[f][j];
            edge_t fin2[4];
            int v2 = 1;
            find_corners(g, B.alive, cells_t, lim_t, B.edge, data, dim, conv, T, line2, liner, fin2, &v2);
            if (v2 && fin2[kAcrossLim[f][j]].slope == 0 && fin[j].slope == 0) {
              expand = 0;
              break;
            }
          }
        }
      }
    }
    if (expand) commit_layer(&B, f, L, data, dim, conv);
  }
  const int np = box_planes(&B, res, origin, points, normals);
  for (int f = 0; f < 6; ++f) free(B.cells[f].v);
  layer_free(L);
  layer_free(T);
  free(line2);
  free(side.v);
  free(tmp_cells.v);
  free(saved);
  return np;
}

/* ---------------------------------------------------------------------------------------------- */
typedef struct cor_params {
  int32_t poly_hor;   /* poly_hor_ */
  int32_t n_it;       /* n_it_decomp_ */
  int32_t rmax;       /* row stride of the polytope arrays */
  int32_t n_traj;     /* points of the previous plan traj_curr_ (N + 1) */
  int32_t max_path;   /* row stride of the path array */
  int32_t reserved;   /* use_cvx_new_: 1 = always GetPolyOcta3DNew (agent_class.cpp:1383) */
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
    int squeezed = 0;
    for (int a = 0; a < 3; ++a) {
      cell_t lo = {{sv[0], sv[1], sv[2]}}, hi = lo;
      lo.c[a] -= 1, hi.c[a] += 1;
      const int in_lo = lo.c[a] >= 0, in_hi = hi.c[a] < dim[a];
      if (in_lo && in_hi && vox(data, dim, lo) == COR_OCC && vox(data, dim, hi) == COR_OCC) squeezed = 1;
    }
    if (squeezed) flags |= COR_FLAG_SQUEEZED;
    double pts[3 * COR_MAX_PLANES], nrm[3 * COR_MAX_PLANES];
    /* a fresh copy of the grid per polytope, like the reference (:1405): the seed voxel is marked even
     * when it is occupied, so marks left behind would read as free space to the next polytope */
    memcpy(work, data, nvox);
    const int np = (squeezed || P->reserved) ? cor_poly_octa_new(sv, work, dim, P->n_it, vs, -(n_poly + 1), origin, pts, nrm)
                                             : cor_poly_octa(sv, work, dim, P->n_it, vs, -(n_poly + 1), origin, pts, nrm);
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
