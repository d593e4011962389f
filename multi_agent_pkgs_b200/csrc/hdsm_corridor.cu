// Safe-corridor generation for sm_100a: one warp per agent, SURVEY.md section 8(f) row 1.
//
// Replaces, per agent and per replanning step:
//   Agent::GenerateSafeCorridor            multi_agent_planner/src/agent_class.cpp:1236-1447
//   convex_decomp_lib::GetPolyOcta3D       convex_decomp_util/src/convex_decomp.cpp:5-376
//   convex_decomp_lib::GetPolyOcta3DNew    convex_decomp_util/src/convex_decomp.cpp:590-1162 (+ FindCorners :378-561)
// and writes the polytope rows straight into the [n][P][Rmax][3] / [n][P][Rmax] / [n][P] arrays that
// hdsm_solve_batch_device consumes (agent_class.cpp:1428-1437), so corridor -> optimisation needs no
// host round trip.
//
// Design.  The decomposition grows a voxel box layer by layer; every step is a short ordered list of
// cells mapped through a few integer tests.  A warp owns one agent: list elements sit on lanes, "does
// any cell fail" is a ballot, "append the survivors in order" a ballot prefix sum, and the few scalar
// decisions (chamfer state machine, limits) are taken by lane 0 in shared memory.  Voxel occupancy is
// never read cell by cell from HBM: per polytope the 32^3 window around the seed is staged once, one
// coalesced 32-byte row per warp load, into a 4 KB bitmap in shared memory ("occupied after
// OccupyUnknown, or outside the grid"); the convex set's own marks (the reference writes CONV into a
// private copy of the grid) are a second 4 KB bitmap.  The window bounds the supported n_it_decomp to
// 90 (15 layers per face; the shipped configurations use 42 and 60).
//
// Integer results (normals, cells) are exact; the few floating-point expressions (hyperplane points,
// b = p . n, the path walk) are written with explicit round-to-nearest intrinsics in the reference's
// evaluation order so that no FMA contraction can change a bit.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <thread>

#include "../../include/hdsm.h"
#include "hdsm_common.h"

namespace hdsm_cor {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kOccVal = 100;   // CVX_DCMP_OCC / ENV_BUILDER_OCC
constexpr int kUnknown = -1;   // ENV_BUILDER_UNK (OccupyUnknown turns it into occupied, voxel_grid.cpp:234-240)
constexpr int kCentre = 16;    // local coordinate of the seed inside the 32^3 window
constexpr int kDq = 128, kDqStart = 32;  // deque storage: <= 31 pushes at either end per layer

// face tables (convex_decomp.cpp:15-43): outward axis/sign, the two in-face growth axes (directions 2, 3
// are their negatives), the box edge met in each in-face direction, the face across it and which of that
// face's limits points along this face's outward direction, and the two faces of every edge.
__constant__ int cOutAxis[6] = {1, 0, 1, 0, 2, 2};
__constant__ int cOutSign[6] = {-1, 1, 1, -1, 1, -1};
__constant__ int cAxA[6][2] = {{0, 2}, {1, 2}, {0, 2}, {1, 2}, {1, 0}, {1, 0}};
__constant__ int cAxS[6][2] = {{1, 1}, {1, 1}, {-1, 1}, {-1, 1}, {-1, 1}, {-1, -1}};
__constant__ int cEdge[6][4] = {{0, 1, 2, 3}, {8, 5, 0, 4}, {10, 9, 8, 11}, {2, 6, 10, 7}, {1, 5, 9, 6}, {3, 7, 11, 4}};
__constant__ int cAcross[6][4] = {{1, 4, 3, 5}, {2, 4, 0, 5}, {3, 4, 1, 5}, {0, 4, 2, 5}, {0, 1, 2, 3}, {0, 3, 2, 1}};
__constant__ int cAcrossLim[6][4] = {{2, 0, 0, 0}, {2, 1, 0, 3}, {2, 2, 0, 2}, {2, 3, 0, 1}, {1, 1, 1, 1}, {3, 3, 3, 3}};
__constant__ int cEdgeFaces[12][2] = {{0, 1}, {0, 4}, {0, 3}, {0, 5}, {1, 5}, {1, 4}, {3, 4}, {3, 5}, {1, 2}, {2, 4}, {2, 3}, {2, 5}};

struct Args {
  hdsm_corridor_params prm;
  int n, cell_cap, layer_cap;
  const int8_t* grids;
  size_t grid_stride;
  const int32_t *grid_index, *dims, *n_path, *prev_n, *prev_rows;
  const double *origins, *pos, *path, *prev_A, *prev_b, *prev_seeds, *prev_traj;
  const uint8_t* prev_used;
  double *poly_A, *poly_b, *seeds;
  int32_t *poly_rows, *flags;
};

// fixed part of the per-warp shared memory; occ[span^2], mark[span^2] (u32), cells[6][cell_cap],
// layer[layer_cap] (u16) and rows[P][Rmax][4] (double) follow
struct Fixed {
  double edge_pos[12][3];
  double et_pos[4][3];
  int edge_slope[12], edge_dir[12], edge_fixed[12], edge_steps[12];
  int et_slope[4], et_dir[4], et_fixed[4], et_steps[4];
  int lim[6][4], alive[6], tip[6], ncell[6];
  int lm[4], ext[4], app[4];
  int top_lo[4], top_hi[4], top_buf[4];  // final state of the four "top" lines of a layer (written once per layer)
  int valid;
  // GetPolyOcta3DNew: limits / edge slopes / line ends of a trial layer (FindCorners), chamfers just started
  int lm2[4], sl2[4], ov_slope[4], ov_dir[4], ov_fixed[4], ov_steps[4];
  int t_lo[4], t_hi[4], t_buf[4], t_ext[4], started[4], fin[4];
  int soft, expand, trial_found;
  unsigned short ring[4][kDq];     // in-face front lines, stored minus (advances so far) * step
  unsigned short top[4][2][kDq];   // their part above the set, double buffered (a failed advance keeps the old one)
  unsigned short ttop[4][2][kDq];  // the same for a trial layer, so that the layer under decision survives
};

__device__ __forceinline__ int coord(unsigned c, int a) { return (c >> (5 * a)) & 31; }
__device__ __forceinline__ unsigned pack(int x, int y, int z) { return (unsigned)(x | (y << 5) | (z << 10)); }
// four 8-bit counters packed in one 32-bit register: uniform over the warp, indexed dynamically with one
// shift and one mask (64-bit variable shifts cost several instructions each)
__device__ __forceinline__ int get8(unsigned v, int q) { return (int)((v >> (8 * q)) & 0xffu); }
__device__ __forceinline__ unsigned set8(unsigned v, int q, int x) { return (v & ~(0xffu << (8 * q))) | ((unsigned)(x & 0xff) << (8 * q)); }

// exact arithmetic helpers: never contracted into FMAs
__device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double dvd(double a, double b) { return __ddiv_rn(a, b); }

struct Agent {
  const Args& A;
  Fixed& S;
  unsigned* occ;          // [span^2] bit x of word (y - wlo) + span (z - wlo): occupied or outside the grid
  unsigned* mark;         // [span^2] voxels of the convex set being grown
  unsigned* posb;         // [span^2] voxel value > 0 after OccupyUnknown, or outside the grid (GetVoxel(...) > 0, :187-209)
  unsigned short* cells;  // [6][cell_cap]
  unsigned short* layer;  // [layer_cap]
  double* rows;           // [P][Rmax][4]: (A0, A1, A2, b) of the polytopes decided so far
  const int lane;
  int dim[3], w0[3];
  int wlo, span;          // staged rows of the window: y, z in [wlo, wlo + span)
  double origin[3];
  const int8_t* grid;
  int flags = 0;

  __device__ Agent(const Args& a, Fixed& s, unsigned char* dyn, int agent) : A(a), S(s), lane(threadIdx.x & 31) {
    const int g = (a.prm.n_it_decomp + 5) / 6 + 1;  // layers a face can gain, plus the cells looked at beyond
    wlo = max(0, kCentre - g);
    span = min(31, kCentre + g) - wlo + 1;
    occ = reinterpret_cast<unsigned*>(dyn);
    mark = occ + span * span;
    posb = mark + span * span;
    cells = reinterpret_cast<unsigned short*>(posb + span * span);
    layer = cells + 6 * a.cell_cap;
    size_t off = (size_t)3 * span * span * sizeof(unsigned) + (size_t)(6 * a.cell_cap + a.layer_cap) * sizeof(unsigned short);
    off = (off + 7) & ~size_t(7);
    rows = reinterpret_cast<double*>(dyn + off);
    for (int k = 0; k < 3; ++k) dim[k] = a.dims[3 * agent + k], origin[k] = a.origins[3 * agent + k];
    grid = a.grids + (size_t)(a.grid_index ? a.grid_index[agent] : agent) * a.grid_stride;
  }

  __device__ __forceinline__ int word_of(unsigned c) const { return (int)((c >> 5) & 31) - wlo + span * ((int)(c >> 10) - wlo); }
  __device__ __forceinline__ bool bit(const unsigned* bm, unsigned c) const { return (bm[word_of(c)] >> (c & 31)) & 1u; }

  // ---------------------------------------------------------------- occupancy window
  __device__ void stage_window(const int seed[3]) {
    for (int k = 0; k < 3; ++k) w0[k] = seed[k] - kCentre;
    for (int i = lane; i < span * span; i += 32) mark[i] = 0u;
    const int gx = w0[0] + lane;
    const bool x_in = gx >= 0 && gx < dim[0];
    for (int r0 = 0; r0 < span * span; r0 += 4) {
      int v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {  // four independent coalesced row loads in flight
        const int r = r0 + u, gy = w0[1] + wlo + r % span, gz = w0[2] + wlo + r / span;
        const bool in = r < span * span && x_in && gy >= 0 && gy < dim[1] && gz >= 0 && gz < dim[2];
        v[u] = in ? (int)__ldg(grid + gx + (size_t)gy * dim[0] + (size_t)gz * dim[0] * dim[1]) : kOccVal;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const unsigned m = __ballot_sync(kFull, v[u] >= kOccVal || v[u] == kUnknown);
        const unsigned m2 = __ballot_sync(kFull, v[u] > 0 || v[u] == kUnknown);
        if (lane == 0 && r0 + u < span * span) occ[r0 + u] = m, posb[r0 + u] = m2;
      }
    }
    __syncwarp();
  }

  // ---------------------------------------------------------------- pieces shared by GetPolyOcta3D / ...New / FindCorners
  struct Lines {  // warp-uniform result of growing the four front lines of one layer
    unsigned TLO, THI, BUF, ext0, ext1, ext2, ext3;
    int nlayer;
    bool overflow;
  };

  __device__ __forceinline__ static int dir_axis(int f, int j) { return cAxA[f][j & 1]; }
  __device__ __forceinline__ static int dir_sign(int f, int j) { return (j < 2 ? 1 : -1) * cAxS[f][j & 1]; }
  __device__ __forceinline__ int along(unsigned c, int f, int j) const {  // cell . in-face direction j, grid coordinates
    const int a = dir_axis(f, j);
    return dir_sign(f, j) * (w0[a] + coord(c, a));
  }

  // limits of the next layer of face f: the face's own (lim4), pulled in where a chamfer is running (:70-91).
  // ovmask bit j: the state of the edge in direction j comes from S.ov_* instead of the box (corners_list_tmp).
  // Lanes 0..3 write out_lm[j]; when `keep` they also copy the edge state into S.et_* (corners_tmp), else only
  // its slope into S.sl2.
  __device__ void layer_limits(int f, const int* lim4, unsigned ovmask, int* out_lm, bool keep) {
    if (lane < 4) {
      const int j = lane, e = cEdge[f][j];
      const bool ov = (ovmask >> j) & 1u;
      int l = lim4[j];
      const int sl = ov ? S.ov_slope[j] : S.edge_slope[e], dr = ov ? S.ov_dir[j] : S.edge_dir[e];
      const int fx = ov ? S.ov_fixed[j] : S.edge_fixed[e], st = ov ? S.ov_steps[j] : S.edge_steps[e];
      if (sl > 0) {
        if (fx) {
          if (dr != f) {
            if (st >= sl) l -= 1;
          } else {
            l -= sl;
          }
        } else if (dr == f) {
          l -= sl;
        }
      }
      out_lm[j] = l;
      if (keep) {
        S.et_slope[j] = sl, S.et_dir[j] = dr, S.et_fixed[j] = fx, S.et_steps[j] = st;
        for (int a = 0; a < 3; ++a) S.et_pos[j][a] = S.edge_pos[e][a];
      } else {
        S.sl2[j] = sl;
      }
    }
    __syncwarp();
  }

  // first cell of the list whose outward neighbour is a free voxel in [1, dim - margin) within the limits
  // (:97-117; margin 1 in GetPolyOcta3D, 0 in GetPolyOcta3DNew and FindCorners); -1 if none
  __device__ int find_seed2d(const unsigned short* cf, int n, int f, const int* lm, int margin) const {
    const int ostep = cOutSign[f] * (1 << (5 * cOutAxis[f]));
    const int a0 = cAxA[f][0], s0 = cAxS[f][0], a1 = cAxA[f][1], s1 = cAxS[f][1];
    const int lm0 = lm[0], lm1 = lm[1], lm2 = lm[2], lm3 = lm[3];
#pragma unroll 1
    for (int base = 0; base < n; base += 32) {
      const int i = base + lane;
      bool ok = false;
      unsigned t = 0;
      if (i < n) {
        t = (unsigned)((int)(cf[i] & 0x7fffu) + ostep);
        const int ta[3] = {w0[0] + coord(t, 0), w0[1] + coord(t, 1), w0[2] + coord(t, 2)};
        ok = ta[0] >= 1 && ta[1] >= 1 && ta[2] >= 1 && ta[0] < dim[0] - margin && ta[1] < dim[1] - margin &&
             ta[2] < dim[2] - margin && !bit(occ, t) && s0 * ta[a0] <= lm0 && s1 * ta[a1] <= lm1 && -s0 * ta[a0] <= lm2 &&
             -s1 * ta[a1] <= lm3;
      }
      const unsigned m = __ballot_sync(kFull, ok);
      if (m) return (int)__shfl_sync(kFull, t, __ffs(m) - 1);
    }
    return -1;
  }

  // In-layer growth from s2: the four front lines advance in turn until none can (:119-209).  A line that
  // failed once can never advance later (its failing cell stays in it; limits, marks and occupancy do not
  // change during the layer), so closed lines are skipped instead of re-tested.  All bookkeeping of the
  // eight deques is warp-uniform register state: RLO / RHI and TLO / THI hold the [lo, hi) bounds of the ring
  // and top lines (one byte per direction), CNT how often a ring line advanced (its cells are stored minus
  // CNT * step, so an advance moves no data), ext0..3 the front of each top line the last time it was
  // non-empty (border_limit_tmp), BUF the live buffer of each top line.  `tp` is the top-line storage
  // ([4][2][kDq]); the cells of the layer are appended to `layer` when lay is set (border_real_tmp).
  __device__ Lines grow_lines(int f, int s2, const int* lm, unsigned short (*tp)[2][kDq], bool lay) {
    const unsigned lt = (1u << lane) - 1u;
    const int ostep = cOutSign[f] * (1 << (5 * cOutAxis[f]));
    if (lane < 4) S.ring[lane][kDqStart] = (unsigned short)s2, tp[lane][0][kDqStart] = (unsigned short)s2;
    if (lay && lane == 0) layer[0] = (unsigned short)s2;
    __syncwarp();
    unsigned RLO = 0x20202020u, RHI = 0x21212121u;  // kDqStart = 32
    Lines L{0x20202020u, 0x21212121u, 0u, (unsigned)s2, (unsigned)s2, (unsigned)s2, (unsigned)s2, 1, false};
    unsigned CNT = 0, open = 15u;
    const auto set_ext = [&](int q, unsigned v) {
      L.ext0 = q == 0 ? v : L.ext0, L.ext1 = q == 1 ? v : L.ext1, L.ext2 = q == 2 ? v : L.ext2, L.ext3 = q == 3 ? v : L.ext3;
    };
#pragma unroll 1
    for (int k = 0; open; ++k) {
      const int j = k & 3;
      if (!((open >> j) & 1u)) continue;
      const int a = dir_axis(f, j), s = dir_sign(f, j);
      const int step = s * (1 << (5 * a));
      const int lo = get8(RLO, j), len = get8(RHI, j) - lo, limit = lm[j];
      const int shift = get8(CNT, j) * step;
      const unsigned short* ring = S.ring[j];
      unsigned short* alt = tp[j][((L.BUF >> j) & 1u) ^ 1u];
      int nr = 0;
      bool ok = true, first_real = false, last_real = false;
      unsigned front_real = 0;
#pragma unroll 1
      for (int base = 0; base < len; base += 32) {
        const int i = base + lane;
        bool fail = false, real = false;
        unsigned t = 0;
        if (i < len) {
          const unsigned c = (unsigned)((int)ring[lo + i] + shift) & 0xffffu;
          if (s * (w0[a] + coord(c, a) + s) > limit) {  // tested before packing: the moved coordinate may leave 0..31
            fail = true;
          } else {
            t = (unsigned)((int)c + step);
            if (bit(mark, (unsigned)((int)t - ostep))) {  // above the set: must be free
              if (bit(occ, t)) fail = true;
              else real = true;
            }
          }
        }
        const unsigned fm = __ballot_sync(kFull, fail), rm = __ballot_sync(kFull, real);
        if (fm) {
          ok = false;
          break;
        }
        if (rm) {
          if (nr == 0) front_real = __shfl_sync(kFull, t, __ffs(rm) - 1);
          if (base == 0) first_real = rm & 1u;
          if (base + 32 >= len) last_real = (rm >> ((len - 1) & 31)) & 1u;
          const int pos = nr + __popc(rm & lt);
          if (real) {
            alt[kDqStart + pos] = (unsigned short)t;
            if (lay && L.nlayer + pos < A.layer_cap) layer[L.nlayer + pos] = (unsigned short)t;
          }
          nr += __popc(rm);
        }
      }
      if (!ok) {
        open &= ~(1u << j);
        continue;
      }
      if (lay && L.nlayer + nr > A.layer_cap) {
        L.overflow = true;
        break;
      }
      // the advance succeeded: ring j moves by one step, its top line is replaced, the neighbours grow
      const unsigned first = (unsigned)((int)ring[lo] + shift + step) & 0xffffu;
      const unsigned last = (unsigned)((int)ring[lo + len - 1] + shift + step) & 0xffffu;
      CNT += 1u << (8 * j);
      L.BUF ^= 1u << j;
      L.nlayer += nr;
      L.TLO = set8(L.TLO, j, kDqStart), L.THI = set8(L.THI, j, kDqStart + nr);
      if (nr > 0) set_ext(j, front_real);
      const int jb = (j + 3) & 3, ja = (j + 1) & 3;
      const int sb = dir_sign(f, jb) * (1 << (5 * dir_axis(f, jb))), sa = dir_sign(f, ja) * (1 << (5 * dir_axis(f, ja)));
      const int hb = get8(RHI, jb), la = get8(RLO, ja) - 1;
      RHI += 1u << (8 * jb), RLO -= 1u << (8 * ja);
      const bool push_tb = nr > 0 && first_real, push_ta = nr > 0 && last_real;
      const int thb = get8(L.THI, jb), tla = get8(L.TLO, ja) - 1;
      if (push_tb) {
        if (thb == get8(L.TLO, jb)) set_ext(jb, first);  // the line was empty: its front changes
        L.THI += 1u << (8 * jb);
      }
      if (push_ta) {
        L.TLO -= 1u << (8 * ja);
        set_ext(ja, last);
      }
      if (lane == 0) {
        S.ring[jb][hb] = (unsigned short)((int)first - get8(CNT, jb) * sb);
        S.ring[ja][la] = (unsigned short)((int)last - get8(CNT, ja) * sa);
        if (push_tb) tp[jb][(L.BUF >> jb) & 1u][thb] = (unsigned short)first;
        if (push_ta) tp[ja][(L.BUF >> ja) & 1u][tla] = (unsigned short)last;
      }
      __syncwarp();
    }
    return L;
  }

  // FindCorners (:378-561): trial of the next layer of face f on the lists cf[0..n) / limits lim4; S.fin[j] = the
  // slope edge j would have afterwards (only "a chamfer starts" is evaluated); returns false when the face is
  // closed or the trial layer would shrink to less than half the area.
  __device__ bool find_corners(int f, const unsigned short* cf, int n, const int* lim4, unsigned ovmask) {
    if (!S.alive[f]) return false;
    layer_limits(f, lim4, ovmask, S.lm2, false);
    if (lane < 4) S.fin[lane] = S.sl2[lane];
    __syncwarp();
    const int s2 = find_seed2d(cf, n, f, S.lm2, 0);
    if (s2 < 0) return true;
    const Lines T = grow_lines(f, s2, S.lm2, S.ttop, false);
    bool ok = true;
    {
      const double area = fabs(fabs((double)S.lm2[0]) - fabs((double)S.lm2[2])) * fabs(fabs((double)S.lm2[1]) - fabs((double)S.lm2[3]));
      const double narea = fabs(fabs((double)along(T.ext0, f, 0)) - fabs((double)along(T.ext2, f, 2))) *
                           fabs(fabs((double)along(T.ext1, f, 1)) - fabs((double)along(T.ext3, f, 3)));
      if (narea < area / 2) ok = false;
    }
    if (lane < 4) {
      const int j = lane, lo = get8(T.TLO, j), hi = get8(T.THI, j);
      if (hi > lo && S.sl2[j] == 0) {
        const int dist = lim4[j] - along(S.ttop[j][(T.BUF >> j) & 1u][lo], f, j);
        if (dist > 0) S.fin[j] = dist;
      }
    }
    __syncwarp();
    return ok;
  }

  // SideIsEmpty (:197-209) over a list: non-empty and every neighbour in direction `istep` holds a value <= 0
  __device__ bool side_is_empty(const unsigned short* v, int n, int istep) const {
    if (n == 0) return false;
    bool any = false;
    for (int i = lane; i < n; i += 32) any |= bit(posb, (unsigned)((int)(v[i] & 0x7fffu) + istep));
    return !__any_sync(kFull, any);
  }

  // ---------------------------------------------------------------- GetPolyOcta3D / GetPolyOcta3DNew
  // Grows the convex set around `seed`; the hyperplanes are written as rows (A = normal, b = point . normal,
  // agent_class.cpp:1428-1437) of polytope slot `slot` in shared memory.  Returns their number, or -1 if a
  // list outgrew its buffer.  use_new selects GetPolyOcta3DNew (convex_decomp.cpp:590-1162).
  __device__ int poly_octa(const int seed[3], int slot, bool use_new) {
    const double res = A.prm.voxel_size;
    const unsigned lt = (1u << lane) - 1u;
    const unsigned sc = pack(kCentre, kCentre, kCentre);
    if (lane == 0) {
      for (int e = 0; e < 12; ++e) S.edge_slope[e] = 0, S.edge_dir[e] = -1, S.edge_fixed[e] = 0, S.edge_steps[e] = 0;
      for (int f = 0; f < 6; ++f) {
        cells[f * A.cell_cap] = (unsigned short)sc;
        S.ncell[f] = 1, S.alive[f] = 1, S.tip[f] = (int)sc;
        const int l0 = cAxS[f][0] * seed[cAxA[f][0]], l1 = cAxS[f][1] * seed[cAxA[f][1]];
        S.lim[f][0] = l0, S.lim[f][1] = l1, S.lim[f][2] = -l0, S.lim[f][3] = -l1;
      }
      mark[word_of(sc)] |= 1u << (sc & 31);
    }
    __syncwarp();
    bool overflow = false;

#pragma unroll 1
    for (int it = 0; it < A.prm.n_it_decomp && !overflow; ++it) {
      const int f = it % 6;
      if (!S.alive[f]) continue;
      const int oa = cOutAxis[f], os = cOutSign[f];
      const int ostep = os * (1 << (5 * oa));
      layer_limits(f, S.lim[f], 0u, S.lm, true);
      const int s2 = find_seed2d(cells + f * A.cell_cap, S.ncell[f], f, S.lm, use_new ? 0 : 1);
      if (s2 < 0) continue;
      const Lines L = grow_lines(f, s2, S.lm, S.top, true);
      if (L.overflow) {
        overflow = true;
        break;
      }
      const int nlayer = L.nlayer;
      if (lane < 4) {
        S.top_lo[lane] = get8(L.TLO, lane), S.top_hi[lane] = get8(L.THI, lane), S.top_buf[lane] = (L.BUF >> lane) & 1u;
        S.ext[lane] = (int)(lane == 0 ? L.ext0 : lane == 1 ? L.ext1 : lane == 2 ? L.ext2 : L.ext3);
      }
      __syncwarp();

      // chamfer bookkeeping of the four edges around the face (:217-301; New: :228-327), lane 0
      if (lane == 0) {
        int valid = 1, soft = 1;
        if (use_new) {  // a layer that shrinks the face to less than half its area is refused (:214-226)
          const double area = fabs(fabs((double)S.lm[0]) - fabs((double)S.lm[2])) * fabs(fabs((double)S.lm[1]) - fabs((double)S.lm[3]));
          const double narea = fabs(fabs((double)along(L.ext0, f, 0)) - fabs((double)along(L.ext2, f, 2))) *
                               fabs(fabs((double)along(L.ext1, f, 1)) - fabs((double)along(L.ext3, f, 3)));
          if (narea < area / 2) soft = 0;
        }
        for (int j = 0; j < 4; ++j) S.started[j] = 0;
        for (int j = 0; j < 4 && valid; ++j) {
          bool stop = false;
          if (S.top_hi[j] > S.top_lo[j]) {
            const unsigned fr = S.top[j][S.top_buf[j]][S.top_lo[j]];
            const int dist = S.lim[f][j] - along(fr, f, j);
            int sl = S.et_slope[j], dr = S.et_dir[j], fx = S.et_fixed[j], st = S.et_steps[j];
            if (sl == 0) {
              if (dist > 0) {
                const int g = cAcross[f][j], ga = cOutAxis[g], gs = cOutSign[g];
                for (int c = 0; c < 3; ++c) {  // cell*res - out*res/2 + out_across*res/2 + res/2 [+ res/2], left to right
                  const double cell = (double)(w0[c] + coord(fr, c));
                  const double o1 = (double)(c == oa ? os : 0), o2 = (double)(c == ga ? gs : 0);
                  double v = add(add(sub(mul(cell, res), dvd(mul(o1, res), 2.0)), dvd(mul(o2, res), 2.0)), dvd(res, 2.0));
                  if (use_new) v = add(v, dvd(res, 2.0));  // (:243-251)
                  S.et_pos[j][c] = v;
                }
                sl = dist, st = dist;
                if (dist > 1) dr = f;
                S.started[j] = dist > 1 ? 2 : 1;
                if (use_new && it < 6) soft = 0;  // no chamfers during the first round (:264-266)
              }
            } else if (fx) {
              if (dr == f || dr == -1) {
                if (dist > sl) {
                  if (use_new) stop = true;  // New leaves the loop here WITHOUT invalidating the layer (:269-271)
                  else valid = 0;
                }
              } else if (st >= sl) {
                if (dist > 1) valid = 0;
                else st = 1;
              } else {
                if (dist != 0) valid = 0;
                else st += 1;
              }
            } else {
              if (dr == -1) {
                if (dist == 0) dr = cAcross[f][j], st += 1, sl += 1;
                else if (dist == 1) fx = 1;
                else valid = 0;
              } else if (dr == f) {
                sl = dist, fx = 1;
              } else {
                if (dist == 0) sl += 1, st += 1;
                else if (dist == 1) fx = 1, st = 1;
                else valid = 0;
              }
            }
            if (stop) break;  // corners_tmp[j..3] keep their copies
            if (valid) S.et_slope[j] = sl, S.et_dir[j] = dr, S.et_fixed[j] = fx, S.et_steps[j] = st;
          }
        }
        S.valid = valid;
        S.soft = soft;
        if (!valid) S.alive[f] = 0;
      }
      __syncwarp();
      if (!S.valid || !S.soft) continue;

      if (use_new) {
        // is there anything to chamfer around?  (:330-366)
        bool expand = true;
        unsigned any_started = 0;
#pragma unroll 1
        for (int j = 0; j < 4; ++j) {
          const int stj = S.started[j];
          if (!stj) continue;
          any_started |= 1u << j;
          const bool first = side_is_empty(S.top[j][S.top_buf[j]] + S.top_lo[j], S.top_hi[j] - S.top_lo[j], ostep);
          bool second = true;
          if (stj == 1) {  // cells of the face across that lie on its limit towards this face
            const int g = cAcross[f][j], q = cAcrossLim[f][j], limg = S.lim[g][q];
            const int gstep = cOutSign[g] * (1 << (5 * cOutAxis[g]));
            const unsigned short* cg = cells + g * A.cell_cap;
            bool have = false, pos = false;
            for (int i = lane; i < S.ncell[g]; i += 32) {
              const unsigned c = cg[i];
              if (along(c, g, q) == limg) have = true, pos |= bit(posb, (unsigned)((int)c + gstep));
            }
            second = __any_sync(kFull, have) && !__any_sync(kFull, pos);
          }
          expand = !(first && second);
          if (!expand) {
            if (lane == 0) S.alive[f] = 0;
            break;
          }
        }
        __syncwarp();
        if (expand && any_started) {
          // trial: mark the layer, look one layer further on this face, unmark (:369-413)
          for (int i = lane; i < nlayer; i += 32) {
            const unsigned c = layer[i];
            const unsigned b = 1u << (c & 31);
            const unsigned old = atomicOr(&mark[word_of(c)], b);
            if (old & b) layer[i] = (unsigned short)(c | 0x8000u);  // was already part of the set
          }
          if (lane < 4) {
            S.t_ext[lane] = along((unsigned)S.ext[lane], f, lane);  // borders_tmp[idx].limits
            S.ov_slope[lane] = S.et_slope[lane], S.ov_dir[lane] = S.et_dir[lane], S.ov_fixed[lane] = S.et_fixed[lane];
            S.ov_steps[lane] = S.et_steps[lane];
          }
          __syncwarp();
          unsigned ovmask = 0;  // corners_list_tmp takes corners_tmp only where no chamfer has just started
          for (int j = 0; j < 4; ++j)
            if (!S.started[j]) ovmask |= 1u << j;
          const bool vfinal = find_corners(f, layer, nlayer, S.t_ext, ovmask);
          int fin[4];
          for (int j = 0; j < 4; ++j) fin[j] = S.fin[j];
          __syncwarp();
          for (int i = lane; i < nlayer; i += 32) {
            const unsigned c = layer[i];
            if (c & 0x8000u) layer[i] = (unsigned short)(c & 0x7fffu);
            else atomicAnd(&mark[word_of(c)], ~(1u << (c & 31)));
          }
          __syncwarp();
          if (vfinal) {
            for (int j = 0; j < 4; ++j)
              if (S.started[j] == 2 && fin[j] < S.et_slope[j]) {
                expand = false;
                if (lane == 0) S.alive[f] = 0;
                break;
              }
            if (expand) {
#pragma unroll 1
              for (int j = 0; j < 4; ++j) {
                if (S.started[j] != 1) continue;
                const int g = cAcross[f][j];
                // borders_tmp[g] equals borders[g] (only face f was replaced); corners_list is the box's own
                const bool v2 = find_corners(g, cells + g * A.cell_cap, S.ncell[g], S.lim[g], 0u);
                const int fin2 = S.fin[cAcrossLim[f][j]];
                __syncwarp();
                if (v2 && fin2 == 0 && fin[j] == 0) {
                  expand = false;
                  break;
                }
              }
            }
          }
          __syncwarp();
        }
        if (!expand) continue;
      }

      // commit the layer (:311-339)
      if (nlayer > A.cell_cap) {
        overflow = true;
        break;
      }
      if (lane == 0) {
        for (int j = 0; j < 4; ++j) {
          const int e = cEdge[f][j];
          S.lim[f][j] = along((unsigned)S.ext[j], f, j);
          S.edge_slope[e] = S.et_slope[j], S.edge_dir[e] = S.et_dir[j], S.edge_fixed[e] = S.et_fixed[j], S.edge_steps[e] = S.et_steps[j];
          for (int c = 0; c < 3; ++c) S.edge_pos[e][c] = S.et_pos[j][c];
          int app = 0;
          if (S.et_slope[j] == 0 && S.top_hi[j] > S.top_lo[j]) {
            const unsigned fr = S.top[j][S.top_buf[j]][S.top_lo[j]];
            if (S.lim[f][j] - along(fr, f, j) == 0) {
              app = 1;  // the face across gains this line of cells and one unit of limit
              S.lim[cAcross[f][j]][cAcrossLim[f][j]] += 1;
            }
          }
          S.app[j] = app;
        }
        S.tip[f] = layer[0];
        S.ncell[f] = nlayer;
      }
      {
        unsigned short* cf = cells + f * A.cell_cap;
        for (int i = lane; i < nlayer; i += 32) {
          const unsigned c = layer[i];
          cf[i] = (unsigned short)c;
          atomicOr(&mark[word_of(c)], 1u << (c & 31));
        }
      }
      __syncwarp();
      for (int j = 0; j < 4; ++j) {
        if (!S.app[j]) continue;
        const int g = cAcross[f][j], lo = S.top_lo[j], cnt = S.top_hi[j] - lo, n0 = S.ncell[g];
        if (n0 + cnt > A.cell_cap) {
          overflow = true;
          break;
        }
        const unsigned short* tp = S.top[j][S.top_buf[j]];
        for (int i = lane; i < cnt; i += 32) cells[g * A.cell_cap + n0 + i] = tp[lo + i];
        __syncwarp();
        if (lane == 0) S.ncell[g] = n0 + cnt;
        __syncwarp();
      }
    }
    if (overflow) return -1;

    // hyperplanes (:343-375): chamfers in edge order, then the six faces; one lane per plane
    const int R = A.prm.max_rows_per_poly;
    const bool cham = lane < 12 && S.edge_slope[lane < 12 ? lane : 0] > 0;
    const unsigned cm = __ballot_sync(kFull, cham);
    const int ncham = __popc(cm), np = ncham + 6;
    if (np <= R && (cham || (lane >= 12 && lane < 18))) {
      double n3[3], p3[3];
      int row;
      if (cham) {
        const int e = lane, sl = S.edge_slope[e], f1 = cEdgeFaces[e][0], f2 = cEdgeFaces[e][1];
        const int steep = S.edge_dir[e] == f1 ? f1 : f2, flat = S.edge_dir[e] == f1 ? f2 : f1;
        for (int c = 0; c < 3; ++c) {
          n3[c] = (double)(sl * (c == cOutAxis[steep] ? cOutSign[steep] : 0) + (c == cOutAxis[flat] ? cOutSign[flat] : 0));
          p3[c] = add(S.edge_pos[e][c], origin[c]);
        }
        row = __popc(cm & lt);
      } else {
        const int f = lane - 12;
        const unsigned tp = (unsigned)S.tip[f];
        for (int c = 0; c < 3; ++c) {
          const double o = (double)(c == cOutAxis[f] ? cOutSign[f] : 0);
          p3[c] = add(add(add(mul((double)(w0[c] + coord(tp, c)), res), dvd(mul(o, res), 2.0)), dvd(res, 2.0)), origin[c]);
          n3[c] = o;
        }
        row = ncham + f;
      }
      double* r = rows + (size_t)(slot * R + row) * 4;
      r[0] = n3[0], r[1] = n3[1], r[2] = n3[2];
      r[3] = add(add(mul(p3[0], n3[0]), mul(p3[1], n3[1])), mul(p3[2], n3[2]));
    }
    __syncwarp();
    return np;
  }

  // LinearConstraint::inside (polyhedron.h:130-137) for polytope slot i: no row with A x - b > 0
  __device__ bool inside(int i, int nrows, const double pt[3]) const {
    bool out = false;
    if (lane < nrows) {
      const double* r = rows + (size_t)(i * A.prm.max_rows_per_poly + lane) * 4;
      out = sub(add(add(mul(r[0], pt[0]), mul(r[1], pt[1])), mul(r[2], pt[2])), r[3]) > 0.0;
    }
    return !__any_sync(kFull, out);
  }

  __device__ void store_poly(int agent, int slot, int nrows, const double sw[3]) {
    const int R = A.prm.max_rows_per_poly, PH = A.prm.poly_hor;
    if (lane < R) {
      const double* r = rows + (size_t)(slot * R + lane) * 4;
      const size_t o = ((size_t)agent * PH + slot) * R + lane;
      const bool on = lane < nrows;
      A.poly_A[o * 3] = on ? r[0] : 0.0, A.poly_A[o * 3 + 1] = on ? r[1] : 0.0, A.poly_A[o * 3 + 2] = on ? r[2] : 0.0;
      A.poly_b[o] = on ? r[3] : 0.0;
    }
    if (lane < 3) A.seeds[((size_t)agent * PH + slot) * 3 + lane] = sw[lane];
    if (lane == 0) A.poly_rows[(size_t)agent * PH + slot] = nrows;
  }

  // ---------------------------------------------------------------- GenerateSafeCorridor
  __device__ void run(int agent) {
    const int R = A.prm.max_rows_per_poly, PH = A.prm.poly_hor;
    int n_poly = 0;
    int nrows_of[HDSM_MAX_POLY];
    double seed_of[HDSM_MAX_POLY][3];
    for (int i = 0; i < HDSM_MAX_POLY; ++i) nrows_of[i] = 0;
    // absent slots read as zero rows
    for (int i = lane; i < PH * R; i += 32) {
      const size_t o = (size_t)agent * PH * R + i;
      A.poly_A[o * 3] = A.poly_A[o * 3 + 1] = A.poly_A[o * 3 + 2] = 0.0, A.poly_b[o] = 0.0;
    }
    if (lane < PH) A.poly_rows[(size_t)agent * PH + lane] = 0;
    if (lane < 3 * PH) A.seeds[(size_t)agent * PH * 3 + lane] = 0.0;
    __syncwarp();

    const int prev_n = A.prev_n ? A.prev_n[agent] : 0;
    const auto keep = [&](int i) {  // previous polytope i becomes slot n_poly
      const int nr = A.prev_rows[(size_t)agent * PH + i];
      if (lane < R) {
        const size_t o = ((size_t)agent * PH + i) * R + lane;
        double* r = rows + (size_t)(n_poly * R + lane) * 4;
        r[0] = A.prev_A[o * 3], r[1] = A.prev_A[o * 3 + 1], r[2] = A.prev_A[o * 3 + 2], r[3] = A.prev_b[o];
      }
      __syncwarp();
      for (int c = 0; c < 3; ++c) seed_of[n_poly][c] = A.prev_seeds[((size_t)agent * PH + i) * 3 + c];
      nrows_of[n_poly] = nr;
      store_poly(agent, n_poly, nr, seed_of[n_poly]);
      ++n_poly;
    };
    if (prev_n > 0) {
      // the whole previous plan inside the last polytope: keep only that one (:1252-1266); otherwise the
      // polytopes the last optimisation used (:1272-1281)
      const int last = prev_n - 1;
      keep(last);
      bool all_in = true;
      for (int j = 0; j < A.prm.n_traj && all_in; ++j) {
        const double* q = A.prev_traj + ((size_t)agent * A.prm.n_traj + j) * 3;
        const double pt[3] = {q[0], q[1], q[2]};
        all_in = inside(0, nrows_of[0], pt);
      }
      if (!all_in) {
        n_poly = 0;
        __syncwarp();
        for (int i = 0; i < prev_n; ++i)
          if (A.prev_used[(size_t)agent * PH + i]) keep(i);
        for (int s = n_poly; s < 1; ++s) {  // slot 0 held the tentative copy of the last polytope
          const double z[3] = {0, 0, 0};
          store_poly(agent, s, 0, z);
        }
      }
    }
    const int n_path = A.n_path[agent];
    if (n_path < 1) {
      if (lane == 0) A.flags[agent] = flags;
      return;
    }
    const double vs = A.prm.voxel_size, samp = dvd(vs, 10.0);
    const double* path = A.path + (size_t)agent * A.prm.max_path * 3;
    double cur[3] = {A.pos[3 * agent], A.pos[3 * agent + 1], A.pos[3 * agent + 2]};
    int path_idx = 1;
#pragma unroll 1
    while (n_poly < PH) {
      const double* nx = path + 3 * (path_idx - 1);
      const double diff[3] = {sub(nx[0], cur[0]), sub(nx[1], cur[1]), sub(nx[2], cur[2])};
      const double dist = __dsqrt_rn(add(add(mul(diff[0], diff[0]), mul(diff[1], diff[1])), mul(diff[2], diff[2])));
      if (dist > samp) {
        for (int c = 0; c < 3; ++c) cur[c] = add(cur[c], dvd(mul(samp, diff[c]), dist));
      } else {
        for (int c = 0; c < 3; ++c) cur[c] = nx[c];
        if (++path_idx == n_path + 1) break;
      }
      bool in_any = false;
      for (int i = 0; i < n_poly && !in_any; ++i) in_any = inside(i, nrows_of[i], cur);
      if (in_any) continue;
      double sp[3] = {cur[0], cur[1], cur[2]};
      if (dist > 0.0) {
        const double m = samp < dist ? samp : dist;  // one sample back: the last point still inside (:1343-1346)
        for (int c = 0; c < 3; ++c) sp[c] = sub(cur[c], dvd(mul(m, diff[c]), dist));
      }
      int sv[3];
      double sw[3];
      for (int c = 0; c < 3; ++c) {
        sv[c] = (int)dvd(sub(sp[c], origin[c]), vs);
        sw[c] = add(add(mul((double)sv[c], vs), dvd(vs, 2.0)), origin[c]);
      }
      bool seen = false;
      for (int i = 0; i < n_poly && !seen; ++i) seen = sw[0] == seed_of[i][0] && sw[1] == seed_of[i][1] && sw[2] == seed_of[i][2];
      if (seen) continue;
      if (sv[0] < 0 || sv[1] < 0 || sv[2] < 0 || sv[0] >= dim[0] || sv[1] >= dim[1] || sv[2] >= dim[2]) {
        flags |= HDSM_COR_SEED_OUTSIDE;
        break;
      }
      stage_window(sv);
      // squeezed seed: the reference switches to GetPolyOcta3DNew (:1385-1395); IsOccupied is false outside the
      // grid, the window bitmap says "occupied" there, hence the explicit range test
      bool squeezed = false;
      for (int c = 0; c < 3; ++c) {
        const bool in_lo = sv[c] - 1 >= 0, in_hi = sv[c] + 1 < dim[c];
        const unsigned ctr = pack(kCentre, kCentre, kCentre);
        if (in_lo && in_hi && bit(occ, ctr - (1u << (5 * c))) && bit(occ, ctr + (1u << (5 * c)))) squeezed = true;
      }
      if (squeezed) flags |= HDSM_COR_SQUEEZED;
      const int np = poly_octa(sv, n_poly, squeezed || A.prm.use_cvx_new != 0);
      if (np < 0) {
        flags |= HDSM_COR_LIST_OVERFLOW;
        break;
      }
      if (np > R) {
        flags |= HDSM_COR_ROW_OVERFLOW;
        break;
      }
      for (int c = 0; c < 3; ++c) seed_of[n_poly][c] = sw[c];
      nrows_of[n_poly] = np;
      store_poly(agent, n_poly, np, sw);
      ++n_poly;
    }
    if (lane == 0) A.flags[agent] = flags;
  }
};

__global__ void __launch_bounds__(32) corridor_kernel(const Args args) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int agent = blockIdx.x;
  if (agent >= args.n) return;
  Fixed& S = *reinterpret_cast<Fixed*>(smem_raw);
  Agent ag(args, S, smem_raw + ((sizeof(Fixed) + 15) & ~size_t(15)), agent);
  ag.run(agent);
}

inline size_t smem_bytes(const hdsm_corridor_params& p, int cell_cap, int layer_cap) {
  const int g = (p.n_it_decomp + 5) / 6 + 1, wlo = kCentre - g > 0 ? kCentre - g : 0;
  const int span = (kCentre + g < 31 ? kCentre + g : 31) - wlo + 1;
  size_t dyn = (size_t)3 * span * span * sizeof(unsigned) + (size_t)(6 * cell_cap + layer_cap) * sizeof(unsigned short);
  dyn = (dyn + 7) & ~size_t(7);
  dyn += (size_t)p.poly_hor * p.max_rows_per_poly * 4 * sizeof(double);
  return ((sizeof(Fixed) + 15) & ~size_t(15)) + dyn;
}

}  // namespace hdsm_cor

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
struct hdsm_corridor {
  hdsm_corridor_params prm{};
  int device = 0, max_agents = 0, max_grids = 0, cell_cap = 0, layer_cap = 0;
  size_t grid_stride = 0, smem = 0;
  cudaStream_t stream = nullptr, stream2 = nullptr;  // stream2: second lane of the chunked host-pointer pipeline
  cudaEvent_t ev_shared = nullptr, ev_chunk[8] = {};
  unsigned char *d_in = nullptr, *h_in = nullptr, *d_out = nullptr, *h_out = nullptr;
  size_t in_cap = 0, out_cap = 0;
  int64_t launches = 0;
  std::string err;
};

namespace {
int cfail(hdsm_corridor* h, int code, const std::string& msg) {
  if (h) h->err = msg;
  return code;
}
#define CCU(call)                                                                              \
  do {                                                                                         \
    cudaError_t e_ = (call);                                                                   \
    if (e_ != cudaSuccess) return cfail(h, HDSM_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
  } while (0)
size_t al256(size_t x) { return (x + 255) & ~size_t(255); }

// memcpy between caller memory and the pinned arenas, large blocks split over a few threads
void staged_copy(void* dst, const void* src, size_t bytes) {
  constexpr size_t kPerThread = size_t(4) << 20;
  const int nt = (int)std::min<size_t>(6, bytes / kPerThread);
  if (nt < 2) {
    memcpy(dst, src, bytes);
    return;
  }
  std::thread th[6];
  const size_t part = ((bytes / nt) + 63) & ~size_t(63);
  for (int t = 1; t < nt; ++t) {
    const size_t o = (size_t)t * part, len = t == nt - 1 ? bytes - o : part;
    th[t] = std::thread([=] { memcpy((char*)dst + o, (const char*)src + o, len); });
  }
  memcpy(dst, src, part);
  for (int t = 1; t < nt; ++t) th[t].join();
}
}  // namespace

extern "C" {

int hdsm_corridor_create(const hdsm_corridor_params* p, int max_agents, int max_grids, size_t grid_stride, int device,
                         hdsm_corridor** out) {
  if (!p || !out || max_agents < 1 || max_grids < 1 || grid_stride < 1) return HDSM_ERR_INVALID;
  if (p->poly_hor < 1 || p->poly_hor > HDSM_MAX_POLY || p->n_it_decomp < 0 || p->n_it_decomp > 90 ||
      p->max_rows_per_poly < 18 || p->max_rows_per_poly > 32 || p->n_traj < 0 || p->max_path < 1 || !(p->voxel_size > 0))
    return HDSM_ERR_INVALID;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return HDSM_ERR_CUDA;  // no CPU fallback
  hdsm_corridor* h = new (std::nothrow) hdsm_corridor();
  if (!h) return HDSM_ERR_INVALID;
  h->prm = *p, h->device = device, h->max_agents = max_agents, h->max_grids = max_grids, h->grid_stride = grid_stride;
  const int g = (p->n_it_decomp + 5) / 6, w = 2 * g + 1;
  h->layer_cap = w * w;              // one face layer
  h->cell_cap = w * w + 4 * g * w;   // plus the lines the four neighbouring faces can hand over
  h->smem = hdsm_cor::smem_bytes(*p, h->cell_cap, h->layer_cap);
  cudaError_t e = cudaSetDevice(device);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->stream2, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_shared, cudaEventDisableTiming);
  for (int c = 0; c < 8 && e == cudaSuccess; ++c) e = cudaEventCreateWithFlags(&h->ev_chunk[c], cudaEventDisableTiming);
  if (e == cudaSuccess) e = hdsm::raise_smem_limit(hdsm_cor::corridor_kernel, device);
  if (e != cudaSuccess) {
    hdsm_corridor_destroy(h);
    return HDSM_ERR_CUDA;
  }
  *out = h;
  return HDSM_OK;
}

void hdsm_corridor_destroy(hdsm_corridor* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->stream2) cudaStreamSynchronize(h->stream2);
  if (h->d_in) cudaFree(h->d_in);
  if (h->h_in) cudaFreeHost(h->h_in);
  if (h->d_out) cudaFree(h->d_out);
  if (h->h_out) cudaFreeHost(h->h_out);
  if (h->ev_shared) cudaEventDestroy(h->ev_shared);
  for (int c = 0; c < 8; ++c)
    if (h->ev_chunk[c]) cudaEventDestroy(h->ev_chunk[c]);
  if (h->stream) cudaStreamDestroy(h->stream);
  if (h->stream2) cudaStreamDestroy(h->stream2);
  delete h;
}

const char* hdsm_corridor_last_error(const hdsm_corridor* h) { return h ? h->err.c_str() : "null handle"; }
int64_t hdsm_corridor_launch_count(const hdsm_corridor* h) { return h ? h->launches : 0; }
int hdsm_corridor_smem_bytes(const hdsm_corridor* h) { return h ? (int)h->smem : 0; }

int hdsm_corridor_batch_device(hdsm_corridor* h, int n, const int8_t* grids, const int32_t* grid_index, const int32_t* dims,
                               const double* origins, const double* pos, const double* path, const int32_t* n_path,
                               const int32_t* prev_n, const double* prev_A, const double* prev_b, const int32_t* prev_rows,
                               const double* prev_seeds, const uint8_t* prev_used, const double* prev_traj, double* poly_A,
                               double* poly_b, int32_t* poly_rows, double* seeds, int32_t* flags, void* stream) {
  if (!h) return HDSM_ERR_INVALID;
  if (n < 0 || !grids || !dims || !origins || !pos || !path || !n_path || !poly_A || !poly_b || !poly_rows || !seeds || !flags)
    return cfail(h, HDSM_ERR_INVALID, "null argument");
  if (prev_n && (!prev_A || !prev_b || !prev_rows || !prev_seeds || !prev_used || (h->prm.n_traj > 0 && !prev_traj)))
    return cfail(h, HDSM_ERR_INVALID, "prev_n given without the previous polytopes");
  if (n > h->max_agents) return cfail(h, HDSM_ERR_CAPACITY, "n exceeds max_agents");
  if (n == 0) return HDSM_OK;
  CCU(cudaSetDevice(h->device));
  hdsm_cor::Args a{};
  a.prm = h->prm, a.n = n, a.cell_cap = h->cell_cap, a.layer_cap = h->layer_cap;
  a.grids = grids, a.grid_stride = h->grid_stride, a.grid_index = grid_index, a.dims = dims, a.n_path = n_path;
  a.prev_n = prev_n, a.prev_rows = prev_rows, a.origins = origins, a.pos = pos, a.path = path, a.prev_A = prev_A;
  a.prev_b = prev_b, a.prev_seeds = prev_seeds, a.prev_traj = prev_traj, a.prev_used = prev_used;
  a.poly_A = poly_A, a.poly_b = poly_b, a.seeds = seeds, a.poly_rows = poly_rows, a.flags = flags;
  cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : h->stream;
  hdsm_cor::corridor_kernel<<<n, 32, h->smem, s>>>(a);
  h->launches += 1;
  CCU(cudaGetLastError());
  return HDSM_OK;
}

int hdsm_corridor_batch(hdsm_corridor* h, int n, int n_grids, const int8_t* grids, const int32_t* grid_index,
                        const int32_t* dims, const double* origins, const double* pos, const double* path,
                        const int32_t* n_path, const int32_t* prev_n, const double* prev_A, const double* prev_b,
                        const int32_t* prev_rows, const double* prev_seeds, const uint8_t* prev_used, const double* prev_traj,
                        double* poly_A, double* poly_b, int32_t* poly_rows, double* seeds, int32_t* flags) {
  if (!h) return HDSM_ERR_INVALID;
  if (n < 0 || n_grids < 1 || !grids || !dims || !origins || !pos || !path || !n_path || !poly_A || !poly_b || !poly_rows ||
      !seeds || !flags)
    return cfail(h, HDSM_ERR_INVALID, "null argument");
  if (n > h->max_agents || n_grids > h->max_grids) return cfail(h, HDSM_ERR_CAPACITY, "n / n_grids exceed the handle's capacity");
  if (!grid_index && n_grids < n) return cfail(h, HDSM_ERR_INVALID, "grid_index is required when agents share grids");
  if (n == 0) return HDSM_OK;
  for (int i = 0; i < n; ++i) {
    const int gi = grid_index ? grid_index[i] : i;
    if (gi < 0 || gi >= n_grids) return cfail(h, HDSM_ERR_INVALID, "grid_index out of range");
    const int32_t* d = dims + 3 * i;
    if (d[0] < 1 || d[1] < 1 || d[2] < 1 || (size_t)d[0] * d[1] * d[2] > h->grid_stride)
      return cfail(h, HDSM_ERR_INVALID, "grid dimensions exceed grid_stride");
    if (n_path[i] < 0 || n_path[i] > h->prm.max_path) return cfail(h, HDSM_ERR_INVALID, "n_path exceeds max_path");
    if (prev_n && (prev_n[i] < 0 || prev_n[i] > h->prm.poly_hor)) return cfail(h, HDSM_ERR_INVALID, "prev_n exceeds poly_hor");
  }
  CCU(cudaSetDevice(h->device));
  const size_t N = (size_t)n, PH = (size_t)h->prm.poly_hor, R = (size_t)h->prm.max_rows_per_poly;
  struct Seg {
    const void* src;
    size_t bytes, off;
  };
  Seg in[16];
  int ni = 0;
  size_t off = 0;
  const auto put = [&](const void* p, size_t bytes) {
    in[ni] = Seg{p, p ? bytes : 0, off};
    if (p) off += al256(bytes);
    return ni++;
  };
  const int i_grid = put(grids, (size_t)n_grids * h->grid_stride), i_gi = put(grid_index, N * 4), i_dim = put(dims, N * 12);
  const int i_org = put(origins, N * 24), i_pos = put(pos, N * 24), i_path = put(path, N * h->prm.max_path * 24);
  const int i_np = put(n_path, N * 4), i_pn = put(prev_n, N * 4);
  const bool pv = prev_n != nullptr;
  const int i_pA = put(pv ? prev_A : nullptr, N * PH * R * 24), i_pb = put(pv ? prev_b : nullptr, N * PH * R * 8);
  const int i_pr = put(pv ? prev_rows : nullptr, N * PH * 4), i_ps = put(pv ? prev_seeds : nullptr, N * PH * 24);
  const int i_pu = put(pv ? prev_used : nullptr, N * PH), i_pt = put(pv ? prev_traj : nullptr, N * h->prm.n_traj * 24);
  if (off > h->in_cap) {
    if (h->d_in) cudaFree(h->d_in);
    if (h->h_in) cudaFreeHost(h->h_in);
    h->d_in = h->h_in = nullptr, h->in_cap = 0;
    CCU(cudaMalloc(&h->d_in, off));
    CCU(cudaMallocHost(&h->h_in, off));
    h->in_cap = off;
  }
  const size_t o_A = 0, o_b = o_A + al256(N * PH * R * 24), o_r = o_b + al256(N * PH * R * 8), o_s = o_r + al256(N * PH * 4);
  const size_t o_f = o_s + al256(N * PH * 24), out_bytes = o_f + al256(N * 4);
  if (out_bytes > h->out_cap) {
    if (h->d_out) cudaFree(h->d_out);
    if (h->h_out) cudaFreeHost(h->h_out);
    h->d_out = h->h_out = nullptr, h->out_cap = 0;
    CCU(cudaMalloc(&h->d_out, out_bytes));
    CCU(cudaMallocHost(&h->h_out, out_bytes));
    h->out_cap = out_bytes;
  }
  // Chunked pipeline over two streams (the voxel grids are 87 KB per agent: the copies cost more than the
  // kernel): chunk c+1 is staged into the pinned arena by a few host threads and copied while chunk c runs;
  // results come back per chunk.  Grids shared through grid_index are staged once, ahead of all chunks.
  const size_t stride_in[16] = {h->grid_stride, 4, 12, 24, 24, (size_t)h->prm.max_path * 24, 4, 4, PH * R * 24, PH * R * 8,
                                PH * 4, PH * 24, PH, (size_t)h->prm.n_traj * 24};  // bytes per agent, in the order of put()
  struct Out {
    void* dst;
    size_t off, stride;
  };
  const Out outs[] = {{poly_A, o_A, PH * R * 24}, {poly_b, o_b, PH * R * 8}, {poly_rows, o_r, PH * 4}, {seeds, o_s, PH * 24}, {flags, o_f, 4}};
  const bool shared_grids = grid_index != nullptr;
  int n_chunks = n >= 2048 ? 4 : 1;
  const int per = (n + n_chunks - 1) / n_chunks;
  if (shared_grids) {
    staged_copy(h->h_in + in[i_grid].off, in[i_grid].src, in[i_grid].bytes);
    CCU(cudaMemcpyAsync(h->d_in + in[i_grid].off, h->h_in + in[i_grid].off, in[i_grid].bytes, cudaMemcpyHostToDevice, h->stream));
  }
  CCU(cudaEventRecord(h->ev_shared, h->stream));
  CCU(cudaStreamWaitEvent(h->stream2, h->ev_shared, 0));
  const auto dp = [&](int k, size_t first) -> const void* {
    if (!in[k].bytes) return nullptr;
    return h->d_in + in[k].off + ((k == i_grid && shared_grids) ? 0 : first * stride_in[k]);
  };
  for (int c = 0; c < n_chunks; ++c) {
    const size_t first = (size_t)c * per;
    if (first >= N) break;
    const size_t cnt = std::min<size_t>(per, N - first);
    cudaStream_t st = (c & 1) ? h->stream2 : h->stream;
    for (int k = 0; k < ni; ++k) {
      if (!in[k].bytes || (k == i_grid && shared_grids)) continue;
      const size_t o = in[k].off + first * stride_in[k], bytes = cnt * stride_in[k];
      staged_copy(h->h_in + o, (const char*)in[k].src + first * stride_in[k], bytes);
      CCU(cudaMemcpyAsync(h->d_in + o, h->h_in + o, bytes, cudaMemcpyHostToDevice, st));
    }
    const int rc = hdsm_corridor_batch_device(
        h, (int)cnt, (const int8_t*)dp(i_grid, first), (const int32_t*)dp(i_gi, first), (const int32_t*)dp(i_dim, first),
        (const double*)dp(i_org, first), (const double*)dp(i_pos, first), (const double*)dp(i_path, first),
        (const int32_t*)dp(i_np, first), (const int32_t*)dp(i_pn, first), (const double*)dp(i_pA, first),
        (const double*)dp(i_pb, first), (const int32_t*)dp(i_pr, first), (const double*)dp(i_ps, first),
        (const uint8_t*)dp(i_pu, first), (const double*)dp(i_pt, first), (double*)(h->d_out + o_A + first * outs[0].stride),
        (double*)(h->d_out + o_b + first * outs[1].stride), (int32_t*)(h->d_out + o_r + first * outs[2].stride),
        (double*)(h->d_out + o_s + first * outs[3].stride), (int32_t*)(h->d_out + o_f + first * outs[4].stride), st);
    if (rc != HDSM_OK) return rc;
    for (const Out& o : outs)
      CCU(cudaMemcpyAsync(h->h_out + o.off + first * o.stride, h->d_out + o.off + first * o.stride, cnt * o.stride,
                          cudaMemcpyDeviceToHost, st));
    CCU(cudaEventRecord(h->ev_chunk[c], st));
  }
  for (int c = 0; c < n_chunks; ++c) {
    const size_t first = (size_t)c * per;
    if (first >= N) break;
    const size_t cnt = std::min<size_t>(per, N - first);
    CCU(cudaEventSynchronize(h->ev_chunk[c]));
    for (const Out& o : outs) memcpy((char*)o.dst + first * o.stride, h->h_out + o.off + first * o.stride, cnt * o.stride);
  }
  return HDSM_OK;
}

}  // extern "C"
