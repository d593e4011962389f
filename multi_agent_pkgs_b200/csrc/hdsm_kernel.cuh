// Batched trajectory optimisation for sm_100a: one agent per thread block of four warps, or per cluster of such blocks.
//
// Replaces, per agent and per replanning step (multi_agent_planner/src/agent_class.cpp):
//   K1  GenerateTimeAwareSafeCorridor  :1086-1215   inter-agent separating planes, built on device
//   K2  SolveOptimizationProblem       :858-1023    + GRBModel::optimize :959  (primal-dual interior point)
//   K3  the binaries b[k][p] / indicator rows :928-940  (exact branch and bound over candidate sets, in rounds of
//       independent nodes dealt to the blocks of a thread-block cluster)
//   K4  PublishTrajectoryFull payload  :645-677     positions packed for the trajectory exchange
//
// Layout: every per-agent quantity lives in shared memory for the whole solve - rows, slacks, multipliers, the
// 3(N-2)-square KKT matrix (the terminal equalities are eliminated by a null-space basis, see hdsm_tables.h) and copies
// of the parameter tables the iterations walk; HBM is touched once for the inputs (coalesced per-agent blocks + the
// neighbour table, which is L2 resident) and once for the outputs.  The KKT matrix is assembled with four threads per
// row, factorised as LDL' with one block barrier per pivot in a sweep that also accumulates the inverse of the unit
// factor, and the two systems of an iteration are solved as products with that inverse (substitution for
// ill-conditioned factors).  FP64 throughout - the reference is double (decomp_basis/data_type.h:50) and cond(K)
// reaches 1e14 near convergence.
#pragma once
#include <cuda_runtime.h>

#include "hdsm_tables.h"

namespace hdsm {

constexpr double kFeasTol = 1e-6;      // Gurobi FeasibilityTol: constant rows are checked, not solved
constexpr double kPruneMargin = 1e-6;  // a row is dropped only if it keeps this slack everywhere reachable
constexpr double kContainTol = 1e-7;   // segment-in-polytope test on a node optimum
constexpr double kPruneRel = 1e-7;     // bound pruning, relative
constexpr double kStepFrac = 0.97;
constexpr double kFastPivot = 1e-6;    // pivots that kept this share of their diagonal: solves through the explicit inverse of the factor
constexpr double kLooseTol = 1e-6;     // accepted at the iteration limit: still inside the 1e-6 KKT target
constexpr int kStackCap = 96;          // >= 1 + N * (P - 1) open nodes (a branching can push up to P sets)
constexpr unsigned kFull = 0xffffffffu;
constexpr int kMaxWidth = 8;           // widest search round = largest portable thread-block cluster
constexpr int kCutoff = 6;             // internal QP status: the dual bound reached the incumbent, solve abandoned
constexpr int kDeferred = 7;           // internal agent status between the passes of one call: search budget of the first pass used up
constexpr int kDeferredOverflow = 8;   // ... and its rows did not fit the row pool of the cluster pass of this tier
#ifndef HDSM_MINBLOCKS
#define HDSM_MINBLOCKS 4  // resident 4-warp blocks per SM the register allocation aims for (128 registers; 5 spills badly)
#endif

struct KernelArgs {
  int n_local, n_rob, rmax, P;
  const int32_t *global_id, *nbr_begin, *nbr_end, *poly_rows, *assign_in;
  const double *x0, *ref, *poly_A, *poly_b, *prev, *all_pos;
  const uint8_t* all_valid;
  const double* bounds;  // optional [n_rob][4]: centre and radius of every plan's points 1..N (radius < 0: no plan), see hdsm_plan_bounds_kernel
  double *traj, *ctrl, *pos_out;
  uint8_t* poly_used;
  int32_t* assign_out;
  hdsm_result* res;
  const int32_t* order;  // optional dispatch order: block b solves agent order[b] (longest expected first)
  int row_cap;     // rows (inter-agent + corridor) one block can hold in shared memory
  unsigned only_mask;   // != 0: solve only agents whose res[].status has its bit set here (later passes of a call)
  int round_budget;     // > 0: an agent whose search is not finished after this many rounds is left to the cluster pass (kDeferred)
  int overflow_status;  // status written when the row pool is too small (HDSM_ROW_OVERFLOW, or kDeferredOverflow inside a cluster pass)
  long long* prof; // optional [n_local][16] cycle counters per phase (HDSM_PROFILE=1), else null
  int max_iter, max_nodes, prune;
  int width;  // nodes per search round (1 = depth-first search); the result depends on it
  int warm_start;  // 1: after the root has branched, the previous plan's assignment is solved in the next round
  int csize;  // thread blocks per agent = cluster size (1, 2 or 4): execution only, never changes a result
  int slot0;  // first dispatch slot of this launch (its blocks cover the slots [slot0, n_local))
  int dbg;  // debugging switches (HDSM_DEBUG): 1 = no dominance filter, 2 = no parent-bound pruning
  double tol;
};

// Shared-memory layout in doubles.  Everything whose size depends only on the horizon sits at a
// compile-time offset (no registers spent on pointers); the polytope and row buffers, sized at run
// time, follow.
struct FixedLayout {
  int Ks, invd, diag0, bs, bl, DQ, TQ, FQ, qv, dq, dqc, qbar, Mk, Tk, Fk, p, dp, dpc, pbar, plo, phi, w, dw, dwc, g, bestw, s0,
      viol, red, sbnd, cbox, tEQ, tQP, tqlo, tqhi, prof, ints, var;
};
HDSM_HD constexpr FixedLayout make_layout(int N) {
  const int NW = 3 * (N - 2), NQ3 = 3 * (3 * N - 2), K3 = 3 * (N + 1);
  FixedLayout s{};
  int o = 0;
  s.Ks = o, o += NW * (NW + 1);
  s.invd = o, o += NW;
  s.diag0 = o, o += NW;
  s.bs = o, o += NQ3 * 2;
  s.bl = o, o += NQ3 * 2;
  s.DQ = o, o += NQ3;
  s.TQ = o, o += NQ3;
  s.FQ = o, o += NQ3;
  s.qv = o, o += NQ3;
  s.dq = o, o += NQ3;
  s.dqc = o, o += NQ3;
  s.qbar = o, o += NQ3;
  s.Mk = o, o += (N + 1) * 6;
  s.Tk = o, o += K3;
  s.Fk = o, o += K3;
  s.p = o, o += K3;
  s.dp = o, o += K3;
  s.dpc = o, o += K3;
  s.pbar = o, o += K3;
  s.plo = o, o += K3;
  s.phi = o, o += K3;
  s.w = o, o += NW;
  s.dw = o, o += NW;
  s.dwc = o, o += NW;
  s.g = o, o += NW;
  s.bestw = o, o += NW;
  s.s0 = o, o += 12;   // x0 (9), c0, spare
  s.red = o, o += 2 * 4 * 4;  // block reductions: 2 buffers x 4 warps x 4 values
  s.viol = o, o += N * kMaxP;
  s.sbnd = o, o += kStackCap;   // parent bound of every open node
  s.cbox = o, o += kMaxP * 6;   // axis-aligned bounding box of every cell: lo[3], hi[3]
  // copies of the parameter tables every interior-point iteration walks (Tables::EQ, QP, qlo, qhi; 8.8 KB at N = 10):
  // read through L1 they cost an L2 round trip whenever the neighbour scans, spills and polytope loads of the four
  // resident blocks have pushed them out, and the assembly loops are chains of such loads
  o += o & 1;  // 16-byte aligned rows
  s.tEQ = o, o += NQ3 * (N - 2);
  s.tQP = o, o += K3 * (N - 2);
  s.tqlo = o, o += NQ3;
  s.tqhi = o, o += NQ3;
  s.prof = o;
#ifdef HDSM_ENABLE_PROFILE
  o += 18;
#endif
  s.ints = o;
  // int32 region: segment tables 4(N+2), bestsig N, fullsig N, prow_n 8, cur 16 B, stack, ctl 8, kp_of_slot N+2, qconst NQ3, pair table u16
  const int int_words = 4 * (N + 2) + 2 * N + kMaxP + 4 + kStackCap * 4 + (NW * (NW + 1) / 2 + 1) / 2 + 2 + 8 + (N + 2) + NQ3;
  o += (int_words + 1) / 2;
  s.var = o;
  return s;
}
// run-time part after FixedLayout::var: poly [P*rmax*4], bmin [P*rmax*P], rown [4][rows] (component-major), rs [rows], rl [rows],
// nid [P*rmax bytes]
// ... then the outcome slots of the search rounds [2][kMaxWidth][6 + (N+1)/2 + 3(N-2)] and the round's node masks
HDSM_HD inline int smem_doubles(int N, int P, int rmax, int row_cap) {
  return make_layout(N).var + P * rmax * 4 + P * rmax * P + row_cap * 6 + (P * rmax + 7) / 8 +
         2 * kMaxWidth * (6 + (N + 1) / 2 + 3 * (N - 2)) + kMaxWidth * 2;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(kFull, v, o));
  return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(kFull, v, o));
  return v;
}
// 1 / d for a normal, positive d: hardware seed (MUFU.RCP64H, ~20 bits) and two Newton steps - four dependent
// multiply-adds, no slow-path call, at most one unit in the last place off the rounded quotient.  The interior-point
// iterations divide only by slacks, multipliers and accepted pivots (all positive and far from the subnormal range).
__device__ __forceinline__ double fast_rcp(double d) {
  double x;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
  double e = fma(-d, x, 1.0);
  x = fma(x, e, x);
  e = fma(-d, x, 1.0);
  return fma(x, e, x);
}

// Separating plane between own point pc and neighbour point po: agent_class.cpp:1152-1205, pert = 0.
// Returns false for coincident points (the reference produces NaN rows there).
__device__ __forceinline__ bool interagent_plane(const hdsm_params& P, const double pc[3], const double po[3], double nf[3],
                                                 double& b) {
  const double nx = po[0] - pc[0], ny = po[1] - pc[1], nz = po[2] - pc[2];
  const double n2 = nx * nx + ny * ny + nz * nz;
  if (!(n2 > 0.0)) return false;
  // one reciprocal square root (1 ulp) instead of a square root and three divisions; the planes are compared with the
  // reference's chain at 1e-13 of the row scale (test_plane_coefficients_match_the_reference_chain)
  const double inrm = rsqrt(n2), nrm = n2 * inrm;
  const double ux = nx * inrm, uy = ny * inrm, uz = nz * inrm;
  // Safety distance of the reference (:1159-1165): ang = pi/2 - |acos(uz)| = asin(uz),
  // t = atan((r/h) tan(ang)), s = hypot(r cos t, h sin t).  With c = uz, rho = r/h:
  // tan(ang) = c / sqrt(1 - c^2), tan(t) = rho c / sqrt(1 - c^2) and
  //     s^2 = (r^2 + h^2 tan^2 t) / (1 + tan^2 t) = r^2 / (1 - c^2 (1 - rho^2)),
  // the same number without the chain of six dependent FP64 transcendentals (agreement with the
  // libm sequence to ~1e-15 relative; tests/test_gpu_parity.py checks the planes through the QPs).
  const double rho = P.drone_radius / P.drone_z_offset, cz = fmin(1.0, fmax(-1.0, uz));
  const double sd = P.drone_radius * rsqrt(1.0 - cz * cz * (1.0 - rho * rho));
  const double h = fmin(2.0 * sd, nrm) / 2.0;
  const double px = (pc[0] + po[0]) / 2 - h * ux, py = (pc[1] + po[1]) / 2 - h * uy, pz = (pc[2] + po[2]) / 2 - h * uz;
  // right = u x (0,0,1) + u x (0,1,0) = (uy - uz, -ux, ux);  up_final = u x (0,1,0) = (-uz, 0, ux)
  nf[0] = P.tilt * (uy - uz) + P.tilt * (-uz) + ux;
  nf[1] = P.tilt * (-ux) + uy;
  nf[2] = P.tilt * ux + P.tilt * ux + uz;
  b = nf[0] * px + nf[1] * py + nf[2] * pz;
  return true;
}

// N: horizon.  W: warps cooperating on one agent (1 or 4).  Set-up, plane assembly and the search
// bookkeeping run on warp 0; the interior-point iterations use all W warps.
template <int N, int W>
struct Solver {
  static constexpr int NZ = N - 2, NW = 3 * NZ, NQ = 3 * N - 2, NQ3 = 3 * NQ, K3 = 3 * (N + 1), LD = NW + 1;
  static constexpr int NT = 32 * W;             // threads per agent
  static constexpr int TPR = W >= 4 ? 4 : 1;    // threads sharing one row of the KKT matrix
  static_assert(W == 1 || W == 4, "W must be 1 or 4");

  const Tables& T;
  const KernelArgs& A;
  const int tid, lane, wid;
  static constexpr FixedLayout L = make_layout(N);
  // shared memory views: fixed-offset arrays are sm + constant, only the last five need registers
  double* const sm;
  double *const Ks = sm + L.Ks, *const invd = sm + L.invd, *const diag0 = sm + L.diag0, *const bs = sm + L.bs,
                *const bl = sm + L.bl, *const DQ = sm + L.DQ, *const TQ = sm + L.TQ, *const FQ = sm + L.FQ,
                *const qv = sm + L.qv, *const dq = sm + L.dq, *const dqc = sm + L.dqc, *const qbar = sm + L.qbar,
                *const Mk = sm + L.Mk, *const Tk = sm + L.Tk, *const Fk = sm + L.Fk, *const p = sm + L.p,
                *const dp = sm + L.dp, *const dpc = sm + L.dpc, *const pbar = sm + L.pbar, *const plo = sm + L.plo,
                *const phi = sm + L.phi, *const w = sm + L.w, *const dw = sm + L.dw, *const dwc = sm + L.dwc,
                *const g = sm + L.g, *const bestw = sm + L.bestw, *const s0 = sm + L.s0, *const viol = sm + L.viol,
                *const red = sm + L.red, *const sbnd = sm + L.sbnd, *const cbox = sm + L.cbox, *const sEQ = sm + L.tEQ,
                *const sQP = sm + L.tQP, *const sqlo = sm + L.tqlo, *const sqhi = sm + L.tqhi;
  int* const ip = reinterpret_cast<int*>(sm + L.ints);
  int *const segb = ip, *const sege = ip + 2 * (N + 2), *const bestsig = ip + 4 * (N + 2), *const fullsig = bestsig + N,
             *const prow_n = fullsig + N;  // segb/sege[2*slot + {0: inter-agent, 1: corridor}]
  unsigned char* const cur = reinterpret_cast<unsigned char*>(prow_n + kMaxP);  // current node's masks [N]
  unsigned char* const stack = cur + 16;  // kStackCap entries of 16 bytes: per-step candidate masks
  int* const ctl = reinterpret_cast<int*>(stack + kStackCap * 16);  // block-wide control words (8)
  int *const skp = ctl + 8, *const sqc = skp + (N + 2);  // Tables::kp_of_slot, Tables::qconst (flat [3 NQ])
  unsigned short* const tab = reinterpret_cast<unsigned short*>(sqc + NQ3);  // pairs (i << 8 | k)
  __device__ __forceinline__ const double* eq_row(int a, int q) const { return sEQ + (a * NQ + q) * NZ; }
  __device__ __forceinline__ const double* qp_row(int a, int k) const { return sQP + (a * (N + 1) + k) * NZ; }
  double *poly, *bmin, *rown, *rs, *rl;  // bmin[id][j]: tightest offset of normal `id` in polytope j (inf: absent)
  unsigned char* nid;  // per polytope row: id of the first row with the same normal
  double c0;
  int nkp, Peff, n_nbr_rows;

  __device__ Solver(const Tables& t, const KernelArgs& a, double* smem)
      : T(t), A(a), tid(threadIdx.x), lane(threadIdx.x & 31), wid(threadIdx.x >> 5), sm(smem) {
    const int rows = a.row_cap;
    poly = sm + L.var;
    bmin = poly + a.P * a.rmax * 4;
    rown = bmin + a.P * a.rmax * a.P;
    rs = rown + rows * 4;
    rl = rs + rows;
    nid = reinterpret_cast<unsigned char*>(rl + rows);
    nkp = T.nkp;
  }

  // ---------------------------------------------------------------- small dense products
  // out[k][a] = (bar ? bar[k][a] : 0) + QP[a][k] . x[a*NZ ...]   (tables from shared memory)
  __device__ __forceinline__ void positions_of(const double* x, const double* bar, double* out) const {
    for (int idx = tid; idx < K3; idx += NT) {
      const int k = idx / 3, a = idx - 3 * k;
      const double* t = qp_row(a, k);
      double v = bar ? bar[idx] : 0.0;
#pragma unroll
      for (int z = 0; z < NZ; ++z) v += t[z] * x[a * NZ + z];
      out[idx] = v;
    }
  }
  __device__ __forceinline__ void quantities_of(const double* x, const double* bar, double* out) const {
    for (int idx = tid; idx < NQ3; idx += NT) {
      const int a = idx / NQ;
      const double* t = sEQ + idx * NZ;
      double v = bar ? bar[idx] : 0.0;
#pragma unroll
      for (int z = 0; z < NZ; ++z) v += t[z] * x[a * NZ + z];
      out[idx] = v;
    }
  }

  // can n.p_kp <= b ever be active inside the reachable box of p_kp?  (exact pruning)
  __device__ __forceinline__ bool reachable(const double n[3], double b, int kp) const {
    double mx = 0;
#pragma unroll
    for (int a = 0; a < 3; ++a) mx += n[a] >= 0 ? n[a] * phi[3 * kp + a] : n[a] * plo[3 * kp + a];
    return !(mx <= b - kPruneMargin);
  }

  // Position rows are processed by groups of `lps` adjacent lanes (a power of two), one group per
  // variable position step, rows strided over the group: the 3x3 barrier blocks reduce inside the
  // group with shuffles.  W = 1: 4 lanes per step at N = 10; W = 4: 16.
  int lps, slot_of_thread, sub;
  __device__ __forceinline__ void init_groups() {
    lps = 32;
    while (lps * nkp > NT) lps >>= 1;
    slot_of_thread = tid / lps, sub = tid % lps;
  }
  template <class F>
  __device__ __forceinline__ void for_my_rows(F&& f) {
    if (slot_of_thread >= nkp) return;
#pragma unroll 1
    for (int sg = 0; sg < 2; ++sg) {
      const int end = sege[2 * slot_of_thread + sg];
#pragma unroll 1
      for (int i = segb[2 * slot_of_thread + sg] + sub; i < end; i += lps) {
        // rows are stored component by component (rown[c][row]): adjacent lanes read adjacent words - with the four
        // components of a row side by side the 32-byte stride cost four shared-memory wavefronts per load
        const double r4[4] = {rown[i], rown[A.row_cap + i], rown[2 * A.row_cap + i], rown[3 * A.row_cap + i]};
        f(r4, rs[i], rl[i]);
      }
    }
  }
  // phase cycle counters: compiled in only with -DHDSM_ENABLE_PROFILE (HDSM_PROFILE=1 then prints them)
#ifdef HDSM_ENABLE_PROFILE
  // counters in shared memory, kept by thread 0 alone: registers would perturb exactly the phases that are short of them
  long long* const tp = reinterpret_cast<long long*>(sm + L.prof);  // [16] phases, [16] time of the last tick
  __device__ __forceinline__ void tick(int slot) {  // attribute the cycles since the last tick to `slot`
    if (A.prof && tid == 0) {
      const long long now = clock64();
      tp[slot] += now - tp[16];
      tp[16] = now;
    }
  }
  __device__ __forceinline__ void tick_start() {
    if (A.prof && tid == 0) tp[16] = clock64();
  }
  __device__ __forceinline__ void tick_init() {
    if (A.prof && tid == 0) {
      for (int i = 0; i < 16; ++i) tp[i] = 0;
      tp[16] = clock64();
    }
  }
#ifdef HDSM_PROF_PASS1  // slots 11-15 split pass 1 instead of the set-up
  __device__ __forceinline__ void tick2(int slot) { tick(slot); }
  __device__ __forceinline__ void tick1(int) {}
#else
  __device__ __forceinline__ void tick2(int) {}
  __device__ __forceinline__ void tick1(int slot) { tick(slot); }
#endif
  __device__ __forceinline__ void tick_flush(int agent) {
    if (A.prof && tid == 0)
      for (int i = 0; i < 16; ++i) A.prof[(size_t)agent * 16 + i] = tp[i];
  }
#else
  __device__ __forceinline__ void tick(int) {}
  __device__ __forceinline__ void tick1(int) {}
  __device__ __forceinline__ void tick2(int) {}
  __device__ __forceinline__ void tick_start() {}
  __device__ __forceinline__ void tick_init() {}
  __device__ __forceinline__ void tick_flush(int) {}
#endif
  __device__ __forceinline__ void bsync() const {
    if (W > 1) __syncthreads();
    else __syncwarp();
  }
  // Reduce four values over the block (bit i of MAXMASK: max instead of sum); result in every thread.
  int red_buf = 0;
  template <int MAXMASK>
  __device__ __forceinline__ void reduce4(double (&v)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = (MAXMASK >> i & 1) ? warp_max(v[i]) : warp_sum(v[i]);
    if (W > 1) {
      double* buf = red + red_buf * 16;
      red_buf ^= 1;
      if (lane == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) buf[wid * 4 + i] = v[i];
      }
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        double r = buf[i];
#pragma unroll
        for (int w2 = 1; w2 < W; ++w2) r = (MAXMASK >> i & 1) ? fmax(r, buf[w2 * 4 + i]) : r + buf[w2 * 4 + i];
        v[i] = r;
      }
    } else {
      __syncwarp();
    }
  }

  // ---------------------------------------------------------------- per-agent set-up
  // block-wide part of the set-up: inputs into shared memory, normal ids, per-polytope offsets, pair table
  __device__ void load_and_index(int agent) {
    if (tid < 9) s0[tid] = A.x0[(size_t)agent * 9 + tid];  // x0 = (p, v, a): s0^a = (x0[a], x0[3+a], x0[6+a])
    for (int t = tid; t < NQ3 * NZ; t += NT) {
      const int aq = t / NZ, z = t - aq * NZ, a = aq / NQ, q = aq - a * NQ;
      sEQ[t] = T.EQ[a][q][z];
    }
    for (int t = tid; t < K3 * NZ; t += NT) {
      const int ak = t / NZ, z = t - ak * NZ, a = ak / (N + 1), k = ak - a * (N + 1);
      sQP[t] = T.QP[a][k][z];
    }
    for (int t = tid; t < NQ3; t += NT) {
      const int a = t / NQ, q = t - a * NQ;
      sqlo[t] = T.qlo[a][q], sqhi[t] = T.qhi[a][q], sqc[t] = T.qconst[a][q];
    }
    if (tid < N + 2) skp[tid] = tid < T.nkp ? T.kp_of_slot[tid] : 0;
    const int PR = A.P * A.rmax;
    for (int i = tid; i < PR; i += NT) {
      const size_t base = (size_t)agent * PR + i;
      poly[4 * i + 0] = A.poly_A[base * 3 + 0];
      poly[4 * i + 1] = A.poly_A[base * 3 + 1];
      poly[4 * i + 2] = A.poly_A[base * 3 + 2];
      poly[4 * i + 3] = A.poly_b[base];
    }
    // clamped: the host entry point validates its arrays, the device entry point cannot
    if (tid < kMaxP) prow_n[tid] = tid < A.P ? min(max(A.poly_rows[(size_t)agent * A.P + tid], 0), A.rmax) : 0;
    // lower-triangle pairs (i, k), k <= i, ordered by row descending: step j of the factorisation touches
    // exactly the pairs with i > j, which are the first NW(NW+1)/2 - (j+1)(j+2)/2 entries
    for (int t = tid; t < NW * NW; t += NT) {
      const int i = t / NW, k = t - i * NW;
      if (k <= i) tab[NW * (NW + 1) / 2 - (i + 1) * (i + 2) / 2 + k] = (unsigned short)((i << 8) | k);
    }
    bsync();
    Peff = 0;
    while (Peff < A.P && prow_n[Peff] > 0) ++Peff;  // P_eff = leading present polytopes (:913)
    // id of a row's normal = flat index of the first valid row (over all polytopes) with a bit-identical
    // normal; 255 marks padding rows.  bmin[id][j] = tightest offset of that normal in polytope j.
    for (int i = tid; i < PR; i += NT) {
      const int j = i / A.rmax, r = i - j * A.rmax;
      int id = 255;
      if (j < Peff && r < prow_n[j]) {
        id = i;
        for (int i2 = 0; i2 < i; ++i2) {
          const int j2 = i2 / A.rmax;
          if (i2 - j2 * A.rmax < prow_n[j2] && poly[4 * i2] == poly[4 * i] && poly[4 * i2 + 1] == poly[4 * i + 1] &&
              poly[4 * i2 + 2] == poly[4 * i + 2]) {
            id = i2;
            break;
          }
        }
      }
      nid[i] = (unsigned char)id;
    }
    bsync();
    for (int t = tid; t < PR * A.P; t += NT) {
      const int i = t / A.P, j = t - i * A.P;
      double bm = INFINITY;
      if (nid[i] == i && j < Peff)
        for (int r2 = 0; r2 < prow_n[j]; ++r2)
          if (nid[j * A.rmax + r2] == i) bm = fmin(bm, poly[4 * (j * A.rmax + r2) + 3]);
      bmin[t] = bm;
    }
    // axis-aligned bounding box of every cell from its rows with a single non-zero component (the six faces
    // GetPolyOcta3D always emits, convex_decomp.cpp:359-373): cbox[6 j + a] = lo_a, cbox[6 j + 3 + a] = hi_a
    for (int t = tid; t < 6 * kMaxP; t += NT) {
      const int j = t / 6, c = t - 6 * j, a = c % 3;
      const bool upper = c >= 3;
      double v = upper ? INFINITY : -INFINITY;
      if (j < Peff)
        for (int r = 0; r < prow_n[j]; ++r) {
          const double* q = poly + 4 * (j * A.rmax + r);
          if ((q[0] != 0.0) + (q[1] != 0.0) + (q[2] != 0.0) != 1 || q[a] == 0.0) continue;
          const double x = q[3] / q[a];
          if (upper && q[a] > 0) v = fmin(v, x);
          if (!upper && q[a] < 0) v = fmax(v, x);
        }
      cbox[t] = v;
    }
    bsync();
  }

  // warp-0 part: objective, reachable boxes, constant checks; returns -1 ok, else an HDSM_* status
  __device__ int setup(int agent) {
    const hdsm_params& P = T.prm;
    const double* ref = A.ref + (size_t)agent * N * 6;
    // gradient and constant of the condensed objective (:870-883, :2098)
    double gi = 0, cpart = 0;
    if (lane < NW) {
      const int a = lane / NZ, r = lane - a * NZ;
      const double sp = s0[a], sv = s0[3 + a], sa = s0[6 + a];
      for (int k = 0; k < N; ++k) {
        const double up = T.Up[a][k][0] * sp + T.Up[a][k][1] * sv + T.Up[a][k][2] * sa;
        gi += 2 * P.r_u * T.Z[a][k][r] * up;
        if (r == 0) cpart += P.r_u * up * up;
      }
      for (int i = 1; i <= N; ++i) {
        const double* wt = i == N ? P.r_n : P.r_x;
        const double ep = T.cP[a][i][0] * sp + T.cP[a][i][1] * sv + T.cP[a][i][2] * sa - ref[(i - 1) * 6 + a];
        const double ev = T.cV[a][i][0] * sp + T.cV[a][i][1] * sv + T.cV[a][i][2] * sa - ref[(i - 1) * 6 + 3 + a];
        gi += 2 * wt[a] * T.QP[a][i][r] * ep + 2 * wt[3 + a] * T.QV[a][i][r] * ev;
        if (r == 0) cpart += wt[a] * ep * ep + wt[3 + a] * ev * ev;
      }
      g[lane] = gi;
    }
    c0 = warp_sum(cpart);
    if (lane == 0) s0[9] = c0;  // the other warps read it from shared memory
    for (int idx = lane; idx < K3; idx += 32) {
      const int k = idx / 3, a = idx - 3 * k;
      pbar[idx] = T.cP[a][k][0] * s0[a] + T.cP[a][k][1] * s0[3 + a] + T.cP[a][k][2] * s0[6 + a];
    }
    for (int idx = lane; idx < NQ3; idx += 32) {
      const int a = idx / NQ, q = idx - a * NQ;
      qbar[idx] = T.cQ[a][q][0] * s0[a] + T.cQ[a][q][1] * s0[3 + a] + T.cQ[a][q][2] * s0[6 + a];
    }
    // reachable box of p_k by interval propagation through the one-step map under the boxes
    if (lane < 3) {
      const int a = lane;
      double lo[3] = {s0[a], s0[3 + a], s0[6 + a]}, hi[3] = {lo[0], lo[1], lo[2]};
      const double alo = a < 2 ? P.min_acc_xy : P.min_acc_z, ahi = a < 2 ? P.max_acc_xy : P.max_acc_z;
      plo[a] = phi[a] = lo[0];
      for (int k = 0; k < N; ++k) {
        double nlo[3], nhi[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          double l = 0, h = 0;
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            const double c = T.A1[a][i][j];
            l += c >= 0 ? c * lo[j] : c * hi[j];
            h += c >= 0 ? c * hi[j] : c * lo[j];
          }
          const double bj = fabs(T.B1[a][i]) * P.max_jerk;
          nlo[i] = l - bj, nhi[i] = h + bj;
        }
        if (k + 1 < N) {  // boxes hold for k = 1..N-1 (:2083-2086)
          nlo[1] = fmax(nlo[1], -P.max_vel), nhi[1] = fmin(nhi[1], P.max_vel);
          nlo[2] = fmax(nlo[2], alo), nhi[2] = fmin(nhi[2], ahi);
          if (nlo[1] > nhi[1]) nlo[1] = nhi[1] = 0.5 * (nlo[1] + nhi[1]);
          if (nlo[2] > nhi[2]) nlo[2] = nhi[2] = 0.5 * (nlo[2] + nhi[2]);
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) lo[i] = nlo[i], hi[i] = nhi[i];
        plo[3 * (k + 1) + a] = lo[0], phi[3 * (k + 1) + a] = hi[0];
      }
    }
    __syncwarp();
    // constant box quantities (v_1 under Euler): feasibility check only
    int bad = 0;
    for (int idx = lane; idx < NQ3; idx += 32) {
      const int a = idx / NQ, q = idx - a * NQ;
      if (T.qconst[a][q] && (qbar[idx] - T.qhi[a][q] > kFeasTol || T.qlo[a][q] - qbar[idx] > kFeasTol)) bad = 1;
    }
    if (__any_sync(kFull, bad)) return HDSM_INFEASIBLE;
    if (Peff == 0) return HDSM_INFEASIBLE;  // sum over an empty set of binaries == 1 (:939-940)
    return -1;
  }

  // ---------------------------------------------------------------- K1: inter-agent rows, packed by position step
  // Quick-reject radius of the plane test for own point ps (previous plan, step k+1) against the
  // reachable box of p_kp: |x - ps| <= rho for every reachable x, the plane keeps distance
  // |n|/2 - smax from ps along u and n_f.u = 1, so the row is inactive if nfmax*rho < |n|/2 - smax - margin.
  __device__ __forceinline__ double reject_thr2(const double ps[3], int kp) const {
    const hdsm_params& P = T.prm;
    const double smax = fmax(P.drone_radius, P.drone_z_offset);
    const double nfmax = 1.0 + 3.0 * fabs(P.tilt);  // |n_f| <= |u| + tilt (|right| + |up|)
    double rho2 = 0;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const double d = fmax(fabs(plo[3 * kp + a] - ps[a]), fabs(phi[3 * kp + a] - ps[a]));
      rho2 += d * d;
    }
    const double thr = 2.0 * (nfmax * sqrt(rho2) + smax + 1e-3);
    return thr * thr;
  }

  // K1 phase A, all warps: compact list of the neighbours that come close enough at any step to give a
  // row that can be active (with 4095 candidates only a few dozen do).  Each warp scans a quarter of
  // the candidate range and appends to its own sub-list (u16 offsets from nbr_begin, kept in the
  // not-yet-used slack array); ctl[2 + w] = entries of warp w, or -1 if its sub-list overflowed.
  __device__ void scan_neighbours(int agent) {
    const int gid = A.global_id[agent];
    const int nb0 = A.nbr_begin ? max(A.nbr_begin[agent], 0) : 0, nb1 = A.nbr_end ? min(A.nbr_end[agent], A.n_rob) : A.n_rob;
    const double* prev = A.prev + (size_t)agent * K3;
    double* thr2k = viol;  // scratch: viol is first written after the first QP
    if (tid < N) {
      const double ps[3] = {prev[3 * (tid + 1)], prev[3 * (tid + 1) + 1], prev[3 * (tid + 1) + 2]};
      thr2k[tid] = fmax(reject_thr2(ps, tid), reject_thr2(ps, tid + 1));
    }
    for (int i = tid; i < K3; i += NT) dp[i] = prev[i];  // own previous positions (dp is scratch here)
    bsync();
    // bounding sphere of the own previous plan (points 1..N) and the largest per-step threshold: a candidate whose
    // plan sphere lies farther than r_own + r_cand + thr_max from it cannot come close at any step, and is skipped
    // on 32 bytes instead of its 24 N bytes of plan (exact: the per-step test below decides everything that remains)
    double cs[3], lim0 = 0;
    if (A.bounds && A.prune) {
      double lo3[3] = {INFINITY, INFINITY, INFINITY}, hi3[3] = {-INFINITY, -INFINITY, -INFINITY}, thrmax2 = 0, rown2 = 0;
      for (int kk = 0; kk < N; ++kk) {
        thrmax2 = fmax(thrmax2, thr2k[kk]);
#pragma unroll
        for (int a = 0; a < 3; ++a) lo3[a] = fmin(lo3[a], dp[3 * kk + 3 + a]), hi3[a] = fmax(hi3[a], dp[3 * kk + 3 + a]);
      }
#pragma unroll
      for (int a = 0; a < 3; ++a) cs[a] = 0.5 * (lo3[a] + hi3[a]);
      for (int kk = 0; kk < N; ++kk) {
        const double dx = dp[3 * kk + 3] - cs[0], dy = dp[3 * kk + 4] - cs[1], dz = dp[3 * kk + 5] - cs[2];
        rown2 = fmax(rown2, dx * dx + dy * dy + dz * dz);
      }
      lim0 = (sqrt(rown2) + sqrt(thrmax2)) * (1.0 + 1e-9) + 1e-9;
    }
    unsigned short* list = reinterpret_cast<unsigned short*>(rs) + wid * A.row_cap;
    const int n = nb1 - nb0, chunk = (n + W - 1) / W;
    const int j_beg = nb0 + wid * chunk, j_end = min(nb1, j_beg + chunk);
    const unsigned lt = (1u << lane) - 1;
    int cnt = 0;
    if (n <= 65535) {
      for (int j0 = j_beg; j0 < j_end; j0 += 32) {
        const int j = j0 + lane;
        bool near = j < j_end && j != gid && A.all_valid[j] != 0;
        if (near && A.prune && A.bounds) {
          const double4 bj = *reinterpret_cast<const double4*>(A.bounds + 4 * (size_t)j);
          const double dx = bj.x - cs[0], dy = bj.y - cs[1], dz = bj.z - cs[2], lim = lim0 + bj.w;
          near = bj.w >= 0.0 && dx * dx + dy * dy + dz * dz <= lim * lim;
        }
        if (near && A.prune) {
          const double* q = A.all_pos + (size_t)j * K3;
          near = false;
#pragma unroll 2
          for (int kk = 0; kk < N; ++kk) {
            const double dx = q[3 * kk + 3] - dp[3 * kk + 3], dy = q[3 * kk + 4] - dp[3 * kk + 4], dz = q[3 * kk + 5] - dp[3 * kk + 5];
            near |= dx * dx + dy * dy + dz * dz <= thr2k[kk];
          }
        }
        const unsigned m = __ballot_sync(kFull, near);
        if (near) {
          const int pos = cnt + __popc(m & lt);
          if (pos < A.row_cap) list[pos] = (unsigned short)(j - nb0);
        }
        cnt += __popc(m);
      }
    } else {
      cnt = A.row_cap + 1;  // offsets do not fit 16 bits: phase B scans the whole range
    }
    if (lane == 0) ctl[2 + wid] = cnt <= A.row_cap ? cnt : -1;
    bsync();
  }

  // K1 phase B, general form on warp 0 (any list length, or no list at all): one plane per (position step, step, neighbour)
  __device__ int build_neighbour_rows_serial(int agent, bool use_list, int n_list) {
    const hdsm_params& P = T.prm;
    const int gid = A.global_id[agent];
    const int nb0 = A.nbr_begin ? max(A.nbr_begin[agent], 0) : 0, nb1 = A.nbr_end ? min(A.nbr_end[agent], A.n_rob) : A.n_rob;
    const double* prev = A.prev + (size_t)agent * K3;
    unsigned short* lists = reinterpret_cast<unsigned short*>(rs);
    int cnt = 0, status = -1;
    const unsigned lt = (1u << lane) - 1;
    for (int kp = 0; kp <= N; ++kp) {
      const int slot = T.slot_of_kp[kp];
      if (slot >= 0) segb[2 * slot] = cnt;
      for (int k = kp - 1; k <= kp; ++k) {
        if (k < 0 || k >= N) continue;
        const double ps[3] = {prev[3 * (k + 1)], prev[3 * (k + 1) + 1], prev[3 * (k + 1) + 2]};
        const double thr2 = reject_thr2(ps, kp);
        const auto process = [&](int j, bool valid) {
          double nf[3] = {0, 0, 0}, b = 0;
          if (valid) {
            const double* q = A.all_pos + ((size_t)j * (N + 1) + k + 1) * 3;
            const double po[3] = {q[0], q[1], q[2]};
            const double dx = po[0] - ps[0], dy = po[1] - ps[1], dz = po[2] - ps[2];
            if (A.prune && dx * dx + dy * dy + dz * dz > thr2) {
              valid = false;
            } else if (!interagent_plane(P, ps, po, nf, b)) {
              status = HDSM_NUMERICAL;
              valid = false;
            } else if (slot < 0) {  // constant point: the row is a number (agent_class.cpp rows on p_0..)
              if (nf[0] * pbar[3 * kp] + nf[1] * pbar[3 * kp + 1] + nf[2] * pbar[3 * kp + 2] - b > kFeasTol)
                status = HDSM_INFEASIBLE;
              valid = false;
            } else if (A.prune && !reachable(nf, b, kp)) {
              valid = false;
            }
          }
          const unsigned m = __ballot_sync(kFull, valid);
          if (valid) {
            const int pos = cnt + __popc(m & lt);
            if (pos < A.row_cap) {
              rown[pos] = nf[0], rown[A.row_cap + pos] = nf[1], rown[2 * A.row_cap + pos] = nf[2], rown[3 * A.row_cap + pos] = b;
            }
          }
          cnt += __popc(m);
        };
        if (use_list) {
          for (int i0 = 0; i0 < n_list; i0 += 32) {
            const int i = i0 + lane;
            process(i < n_list ? nb0 + lists[i] : 0, i < n_list);
          }
        } else {
          for (int j0 = nb0; j0 < nb1; j0 += 32) {
            const int j = j0 + lane;
            process(j, j < nb1 && j != gid && A.all_valid[j] != 0);
          }
        }
      }
      if (slot >= 0) sege[2 * slot] = cnt;
    }
    n_nbr_rows = cnt;
    status = __reduce_max_sync(kFull, status);
    if (cnt > A.row_cap) return HDSM_ROW_OVERFLOW;
    __syncwarp();
    return status;
  }

  // The plane of (step k, neighbour j) and what becomes of it on the two points it acts on: bit 0 / bit 1 of `use` =
  // it is a row on p_k / on p_k+1 (close enough - thr0 / thr1 are reject_thr2 of the two points - and active somewhere
  // in the reachable box); st = -1 or the status it forces (coincident points: NUMERICAL; violated at a constant
  // point: INFEASIBLE).  Static and not inlined: three call sites, and a member function would pin the whole solver
  // object in local memory.
  struct PlaneEval {
    double n0, n1, n2, b;
    int use, st;
  };
  __device__ __noinline__ static void eval_plane(const Tables& T, const KernelArgs& A, const double* sm, double ps0, double ps1,
                                                 double ps2, double thr0, double thr1, int j, int k, PlaneEval& e) {
    e.n0 = e.n1 = e.n2 = e.b = 0.0, e.use = 0, e.st = -1;
    if (j < 0) return;
    const double *plo = sm + L.plo, *phi = sm + L.phi, *pbar = sm + L.pbar;
    const double ps[3] = {ps0, ps1, ps2};
    const double* q = A.all_pos + ((size_t)j * (N + 1) + k + 1) * 3;
    const double po[3] = {q[0], q[1], q[2]};
    const double dx = po[0] - ps[0], dy = po[1] - ps[1], dz = po[2] - ps[2], d2 = dx * dx + dy * dy + dz * dz;
    const bool v[2] = {!A.prune || !(d2 > thr0), !A.prune || !(d2 > thr1)};
    if (!v[0] && !v[1]) return;
    double nf[3], b;
    if (!interagent_plane(T.prm, ps, po, nf, b)) {
      e.st = HDSM_NUMERICAL;
      return;
    }
    e.n0 = nf[0], e.n1 = nf[1], e.n2 = nf[2], e.b = b;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (!v[h]) continue;
      const int kp = k + h;
      if (T.slot_of_kp[kp] < 0) {  // constant point: the row is a number (agent_class.cpp rows on p_0..)
        if (nf[0] * pbar[3 * kp] + nf[1] * pbar[3 * kp + 1] + nf[2] * pbar[3 * kp + 2] - b > kFeasTol) e.st = max(e.st, (int)HDSM_INFEASIBLE);
      } else {
        bool act = true;
        if (A.prune) {  // reachable(): can the row be active inside the reachable box of p_kp?
          double mx = 0;
#pragma unroll
          for (int a = 0; a < 3; ++a) mx += nf[a] >= 0 ? nf[a] * phi[3 * kp + a] : nf[a] * plo[3 * kp + a];
          act = !(mx <= b - kPruneMargin);
        }
        if (act) e.use |= 1 << h;
      }
    }
  }

  // K1 phase B, all warps: planes of the listed neighbours, packed by position step.  The plane of (step k, neighbour)
  // acts on p_k and on p_k+1 and is computed once for both (round 2 computed it for either point), by the warp that
  // owns step k (k = w, w + W, ...).  Two sweeps around one block barrier: the first counts the rows of every
  // (step, point) pair, the second writes them where the segment order (kp ascending; inside kp the rows of step
  // kp - 1, then those of step kp; neighbours in list order - the order the serial form produces) puts them.  Lists of
  // up to 64 neighbours keep their planes in registers between the sweeps, longer ones are evaluated again.  Without
  // a list (a sub-list overflowed) the serial form runs on warp 0.  Returns -1 or an HDSM_* status, uniform over the block.
  __device__ int build_neighbour_rows(int agent) {
    const int nb0 = A.nbr_begin ? max(A.nbr_begin[agent], 0) : 0;
    const double* prev = A.prev + (size_t)agent * K3;
    unsigned short* lists = reinterpret_cast<unsigned short*>(rs);
    int* cnts = reinterpret_cast<int*>(stack);  // [2N] scratch: the search stack is not in use yet
    if (wid == 0) {
      bool use_list = true;
      for (int w2 = 0; w2 < W; ++w2) use_list &= ctl[2 + w2] >= 0;
      // The W sub-lists are packed into one: every round below costs a full plane computation for all 32 lanes,
      // and with a dozen neighbours four quarter-full rounds per step are four times the work of one.
      int n_list = 0;
      if (use_list) {
        n_list = ctl[2];
        for (int w2 = 1; w2 < W; ++w2) {
          const int nl = ctl[2 + w2];
          for (int i0 = 0; i0 < nl; i0 += 32) {  // destination never lies behind the source: read, then write
            const int i = i0 + lane;
            const unsigned short v = i < nl ? lists[w2 * A.row_cap + i] : (unsigned short)0;
            __syncwarp();
            if (i < nl) lists[n_list + i] = v;
            __syncwarp();
          }
          n_list += nl;
        }
      }
      int serial = -2;
      if (W == 1 || !use_list) serial = build_neighbour_rows_serial(agent, use_list, n_list);
      if (lane == 0) ctl[6] = n_list, ctl[7] = serial;
    }
    bsync();
    if (W == 1 || ctl[7] != -2) return ctl[7];
    const int n_list = ctl[6];
    const bool cached = n_list <= 64;
    constexpr int KPW = W == 1 ? 1 : (N + W - 1) / W;  // steps per warp
    double cn[KPW][2][4];
    unsigned mk[KPW][2];  // per cached round: ballots of bit 0 (low half) and bit 1 (high half) of `use`
    int status = -1;
#pragma unroll
    for (int t = 0; t < KPW; ++t) {
      const int k = wid + t * W, kc = min(k, N - 1);
      const double ps[3] = {prev[3 * (kc + 1)], prev[3 * (kc + 1) + 1], prev[3 * (kc + 1) + 2]};
      const double thr0 = reject_thr2(ps, kc), thr1 = reject_thr2(ps, kc + 1);
      int c0 = 0, c1 = 0;
      PlaneEval e;
      if (cached) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const int i = r * 32 + lane;
          eval_plane(T, A, sm, ps[0], ps[1], ps[2], thr0, thr1, (k < N && i < n_list) ? nb0 + lists[i] : -1, kc, e);
          status = max(status, e.st);
          cn[t][r][0] = e.n0, cn[t][r][1] = e.n1, cn[t][r][2] = e.n2, cn[t][r][3] = e.b;
          const unsigned m0 = __ballot_sync(kFull, e.use & 1), m1 = __ballot_sync(kFull, e.use & 2);
          mk[t][r] = (m0 >> lane & 1) | ((m1 >> lane & 1) << 1);  // own bits are enough: the sweeps recount by ballot
          c0 += __popc(m0), c1 += __popc(m1);
        }
      } else if (k < N) {
        for (int i0 = 0; i0 < n_list; i0 += 32) {
          const int i = i0 + lane;
          eval_plane(T, A, sm, ps[0], ps[1], ps[2], thr0, thr1, i < n_list ? nb0 + lists[i] : -1, k, e);
          status = max(status, e.st);
          c0 += __popc(__ballot_sync(kFull, e.use & 1)), c1 += __popc(__ballot_sync(kFull, e.use & 2));
        }
      }
      if (k < N && lane == 0) cnts[2 * k] = c0, cnts[2 * k + 1] = c1;
    }
    status = __reduce_max_sync(kFull, status);
    if (lane == 0) ctl[2 + wid] = status;
    bsync();
    // segment f = 2 k + h starts at the sum of the counts before it; position step kp owns f = 2 kp - 1 and 2 kp
    if (wid == 0) {
      int run = 0;
      for (int f = 0; f < 2 * N; ++f) {
        const int c = cnts[f];
        const int kp_b = (f + 1) / 2;  // f opens the segment of kp_b when f = 2 kp_b - 1 (or f = 0)
        if ((f == 0 || (f & 1)) && lane == 0 && T.slot_of_kp[kp_b] >= 0) segb[2 * T.slot_of_kp[kp_b]] = run;
        run += c;
        const int kp_e = f / 2;        // f closes the segment of kp_e when f = 2 kp_e (or f = 2 N - 1: kp = N)
        if (!(f & 1) && lane == 0 && T.slot_of_kp[kp_e] >= 0) sege[2 * T.slot_of_kp[kp_e]] = run;
        if (f == 2 * N - 1 && lane == 0 && T.slot_of_kp[N] >= 0) sege[2 * T.slot_of_kp[N]] = run;
      }
      if (lane == 0) ctl[6] = run;
    }
    const unsigned lt = (1u << lane) - 1;
    const auto put = [&](int pos, double n0, double n1, double n2, double b) {
      if (pos < A.row_cap) rown[pos] = n0, rown[A.row_cap + pos] = n1, rown[2 * A.row_cap + pos] = n2, rown[3 * A.row_cap + pos] = b;
    };
#pragma unroll
    for (int t = 0; t < KPW; ++t) {
      const int k = wid + t * W;
      if (k >= N) continue;
      int base0 = 0;
      for (int f = 0; f < 2 * k; ++f) base0 += cnts[f];
      int base1 = base0 + cnts[2 * k];
      if (cached) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const unsigned m0 = __ballot_sync(kFull, mk[t][r] & 1), m1 = __ballot_sync(kFull, mk[t][r] & 2);
          if (mk[t][r] & 1) put(base0 + __popc(m0 & lt), cn[t][r][0], cn[t][r][1], cn[t][r][2], cn[t][r][3]);
          if (mk[t][r] & 2) put(base1 + __popc(m1 & lt), cn[t][r][0], cn[t][r][1], cn[t][r][2], cn[t][r][3]);
          base0 += __popc(m0), base1 += __popc(m1);
        }
      } else {
        const double ps[3] = {prev[3 * (k + 1)], prev[3 * (k + 1) + 1], prev[3 * (k + 1) + 2]};
        const double thr0 = reject_thr2(ps, k), thr1 = reject_thr2(ps, k + 1);
        PlaneEval e;
        for (int i0 = 0; i0 < n_list; i0 += 32) {
          const int i = i0 + lane;
          eval_plane(T, A, sm, ps[0], ps[1], ps[2], thr0, thr1, i < n_list ? nb0 + lists[i] : -1, k, e);
          const unsigned m0 = __ballot_sync(kFull, e.use & 1), m1 = __ballot_sync(kFull, e.use & 2);
          if (e.use & 1) put(base0 + __popc(m0 & lt), e.n0, e.n1, e.n2, e.b);
          if (e.use & 2) put(base1 + __popc(m1 & lt), e.n0, e.n1, e.n2, e.b);
          base0 += __popc(m0), base1 += __popc(m1);
        }
      }
    }
    int st = -1;
    for (int w2 = 0; w2 < W; ++w2) st = max(st, ctl[2 + w2]);
    bsync();
    n_nbr_rows = ctl[6];
    if (n_nbr_rows > A.row_cap) return HDSM_ROW_OVERFLOW;
    return st;
  }

  // candidate sets of the root node: polytopes whose rows hold at the constant points (p_0, p_1, ...)
  __device__ bool root_sets(int agent) {
    bool any_empty = false;
    for (int k = 0; k < N; ++k) {
      unsigned mask = 0;
      const int forced = A.assign_in ? A.assign_in[(size_t)agent * N + k] : -1;
      for (int j = 0; j < Peff; ++j) {
        if (forced >= 0 && forced != j) continue;
        int bad = 0;
        for (int kp = k; kp <= k + 1; ++kp) {
          if (T.slot_of_kp[kp] >= 0) continue;
          if (lane < prow_n[j]) {
            const double* r = poly + 4 * (j * A.rmax + lane);
            if (r[0] * pbar[3 * kp] + r[1] * pbar[3 * kp + 1] + r[2] * pbar[3 * kp + 2] - r[3] > kFeasTol) bad = 1;
          }
        }
        if (!__any_sync(kFull, bad)) mask |= 1u << j;
      }
      if (lane == 0) cur[k] = (unsigned char)mask;
      any_empty |= mask == 0;
    }
    __syncwarp();
    return !any_empty;
  }

  // Per-step dominance between cells.  Cell A dominates cell B at step k when every point the segment
  // (p_k, p_k+1) can reach inside B also lies in A: a trajectory that uses B at step k may use A instead at the
  // same cost, so B leaves the candidate set (the optimum value is unchanged; near-duplicate overlapping cells
  // otherwise make the search enumerate assignments that tie).  Sufficient test per row (n, b) of A and per
  // variable point kp in {k, k+1}: B holds the same normal at least as tight, or the row cannot be violated
  // inside reach_box(kp) /\ bbox(B).  One row of A per lane.
  __device__ __forceinline__ bool dominates(int k, int ca, int cb) const {
    bool fail = false;
    if (lane < prow_n[ca]) {
      const int i = ca * A.rmax + lane;
      const double* r = poly + 4 * i;
      const double same = bmin[nid[i] * A.P + cb];
      if (!(same <= r[3])) {
        const double *lo = cbox + 6 * cb, *hi = lo + 3;
        for (int kp = k; kp <= k + 1; ++kp) {
          if (T.slot_of_kp[kp] < 0) continue;  // a constant point was checked against both cells already
          double mx = 0;
#pragma unroll
          for (int a = 0; a < 3; ++a)
            mx += r[a] >= 0 ? r[a] * fmin(phi[3 * kp + a], hi[a]) : r[a] * fmax(plo[3 * kp + a], lo[a]);
          if (!(mx <= r[3])) fail = true;
        }
      }
    }
    return !__any_sync(kFull, fail);
  }
  // all warps, one step per warp and round: drop dominated cells from the root sets (equal cells: the lower index stays)
  __device__ void dominance_filter() {
    for (int k = wid; k < N; k += W) {
      unsigned mask = cur[k];
      if (__popc(mask) > 1) {
        for (int cb = Peff - 1; cb >= 0; --cb) {
          if (!(mask >> cb & 1)) continue;
          for (int ca = 0; ca < Peff; ++ca) {
            if (ca == cb || !(mask >> ca & 1)) continue;
            if (!dominates(k, ca, cb)) continue;
            if (ca > cb && dominates(k, cb, ca)) continue;
            mask &= ~(1u << cb);
            break;
          }
        }
        if (lane == 0) cur[k] = (unsigned char)mask;
      }
    }
    bsync();
  }

  // Row of a candidate set for the distinct normal whose first occurrence is flat row i: offset = max
  // over the members of their tightest offset for that normal; absent in any member -> no row.  For a
  // single polytope this is the polytope itself (duplicate normals collapse to the tightest one); for
  // several it is the union hull used by the relaxation.
  __device__ __forceinline__ bool hull_row(unsigned mask, int i, double n[3], double& b) const {
    if (i >= A.P * A.rmax || nid[i] != i) return false;
    double bmax = -INFINITY;
#pragma unroll 1
    for (int j = 0; j < Peff; ++j)
      if (mask >> j & 1) bmax = fmax(bmax, bmin[i * A.P + j]);
    if (!(bmax < INFINITY)) return false;
    n[0] = poly[4 * i], n[1] = poly[4 * i + 1], n[2] = poly[4 * i + 2], b = bmax;
    return true;
  }

  // static corridor rows of the current node, packed by position step; returns row count or -1 on overflow
  __device__ int build_static_rows() {
    int cnt = 0;
    const int stat_cap = A.row_cap - n_nbr_rows;
    const unsigned lt = (1u << lane) - 1;
    const int rounds = (A.P * A.rmax + 31) / 32;
    for (int slot = 0; slot < nkp; ++slot) {
      const int kp = T.kp_of_slot[slot];
      segb[2 * slot + 1] = n_nbr_rows + cnt;
      for (int k = kp - 1; k <= kp; ++k) {
        if (k < 0 || k >= N) continue;
        if (k == kp && kp >= 1 && cur[kp - 1] == cur[kp]) continue;  // same rows already put on p_kp by step kp-1
        for (int rd = 0; rd < rounds; ++rd) {
          double n[3], b;
          bool valid = hull_row(cur[k], rd * 32 + lane, n, b);
          if (valid && A.prune) valid = reachable(n, b, kp);
          const unsigned m = __ballot_sync(kFull, valid);
          if (valid) {
            const int pos = cnt + __popc(m & lt);
            if (pos < stat_cap) {
              double* r = rown + n_nbr_rows + pos;
              r[0] = n[0], r[A.row_cap] = n[1], r[2 * A.row_cap] = n[2], r[3 * A.row_cap] = b;
            }
          }
          cnt += __popc(m);
        }
      }
      sege[2 * slot + 1] = n_nbr_rows + min(cnt, stat_cap);
    }
    __syncwarp();
    return cnt > stat_cap ? -1 : cnt;
  }

  // ---------------------------------------------------------------- K2: Mehrotra predictor-corrector
  // box row helpers: side 0 = upper (q <= hi), side 1 = lower (lo <= q); coefficient sign +1 / -1
  __device__ __forceinline__ double box_slack(int idx, int side, int a, int q) const {
    return side == 0 ? sqhi[idx] - qv[idx] : qv[idx] - sqlo[idx];
  }

  // LDL^T of Ks (lower part, in place), all threads of the block, one barrier per pivot - and, in the same
  // sweep, the INVERSE of the unit factor, accumulated in the upper triangle, which the factorisation does not use:
  // entry (k, i), k < i, ends up holding Linv[i][k].  Column j is kept unscaled (c_ij = L_ij d_j).  At step j the
  // pair (i, k), k <= i, j < i, does exactly one of
  //     j < k :  K[i][k]    -= c_ij c_kj / d_j            trailing matrix
  //     j = k :  Linv[i][k]  = -c_ij / d_j                forward elimination applied to the identity
  //     j > k :  Linv[i][k] -= (c_ij / d_j) Linv[j][k]
  // and c_kj (lower triangle, j < k) and Linv[j][k] (upper triangle, j > k) are the SAME shared-memory word
  // Ks[k][j]: one predicated multiply-add per pair and step whatever the case, so the inverse rides along at no
  // extra latency (each thread holds its <= PM pairs in registers; one warp: pairs from the table).  With it the
  // two triangular solves - 46 dependent shuffle + multiply-add steps on one warp - become two products that all
  // threads share (solve_inplace).  A pivot that lost all its digits to cancellation (<= 1e-13 of the original
  // diagonal) marks a direction the barrier function has pinned: the variable is frozen for this solve (1/d := 0)
  // instead of aborting.  Only a non-positive / NaN original diagonal is a failure.
  // Tried and measured slower on B200 (profiles/r3_experiments.md): skipping finished pair slots per warp (more
  // branches than it saves), and the whole factorisation on one warp with the matrix in registers and columns by
  // shuffle (5 k instructions of straight-line code per call: instruction fetch bound, 26 k cycles); every pair's
  // running value in a register, stored only when another thread needs it (half the shared-memory traffic, but
  // 226 k instead of 125 k cycles per agent in the factorisation: the conditional stores and selects cost more).
  static constexpr int PM = W == 1 ? 0 : (NW * (NW + 1) / 2 + NT - 1) / NT;
  int pr_ik[PM > 0 ? PM : 1], pr_i[PM > 0 ? PM : 1], pr_k[PM > 0 ? PM : 1];
  __device__ __forceinline__ void init_pairs() {
#pragma unroll
    for (int m = 0; m < PM; ++m) {
      const int t = tid + m * NT;
      const int ik = t < NW * (NW + 1) / 2 ? tab[t] : 0, i = ik >> 8, kk = ik & 255;  // padding: pair (0, 0), never active
      pr_ik[m] = ik, pr_i[m] = i * LD, pr_k[m] = kk * LD;
    }
  }
  __device__ __forceinline__ void pair_step(int j, double inv, int i, int kk, int oi, int ok) {
    const double l = Ks[oi + j] * inv;
    if (j < kk) {
      Ks[oi + kk] -= l * Ks[ok + j];
    } else if (kk < i) {  // (a diagonal pair is finished once j reaches it)
      const double old = j == kk ? 0.0 : Ks[ok + i], b = j == kk ? 1.0 : Ks[ok + j];
      Ks[ok + i] = old - l * b;
    }
  }
  bool fast_solve = true;  // set by factor(): no accepted pivot lost more than six digits (block-uniform)
  __device__ __forceinline__ bool factor() {
    bool ok = true;
    fast_solve = true;
#pragma unroll 1
    for (int j = 0; j < NW - 1; ++j) {
      if (W == 1) {
        const double djj = Ks[j * LD + j], dor = diag0[j];
        ok &= dor > 0.0;
        const double inv = djj > 1e-13 * dor ? fast_rcp(djj) : 0.0;
        fast_solve &= !(djj > 1e-13 * dor) || djj >= kFastPivot * dor;
        if (tid == j) invd[j] = inv;
        const int T_j = NW * (NW + 1) / 2 - (j + 1) * (j + 2) / 2;
#pragma unroll 3
        for (int t = tid; t < T_j; t += NT) {
          const int ik = tab[t], i = ik >> 8, kk = ik & 255;
          pair_step(j, inv, i, kk, i * LD, kk * LD);
        }
      } else {
        // all loads of the thread's pairs in flight before the pivot's reciprocal is needed and before the first
        // store: one shared-memory round trip per pivot instead of one per pair
        constexpr int PA = PM > 0 ? PM : 1;
        double a[PA], b[PA], old[PA];
        int dst[PA];
#pragma unroll
        for (int m = 0; m < PM; ++m) {
          const int i = pr_ik[m] >> 8, kk = pr_ik[m] & 255;
          dst[m] = j < kk ? pr_i[m] + kk : pr_k[m] + i;
          a[m] = b[m] = old[m] = 0.0;
          if (j < i) a[m] = Ks[pr_i[m] + j], b[m] = Ks[pr_k[m] + j], old[m] = Ks[dst[m]];  // predicated: finished pairs cost no shared-memory bandwidth
        }
        const double djj = Ks[j * LD + j], dor = diag0[j];
        ok &= dor > 0.0;
        const double inv = djj > 1e-13 * dor ? fast_rcp(djj) : 0.0;
        fast_solve &= !(djj > 1e-13 * dor) || djj >= kFastPivot * dor;
        if (tid == j) invd[j] = inv;
#pragma unroll
        for (int m = 0; m < PM; ++m) {
          const int i = pr_ik[m] >> 8, kk = pr_ik[m] & 255;
          const double bb = j == kk ? 1.0 : b[m], oo = j == kk ? 0.0 : old[m];
          const double r = oo - (a[m] * inv) * bb;
          if (j < i) Ks[dst[m]] = r;
        }
      }
      bsync();
    }
    {  // last pivot: nothing left to update
      const double djj = Ks[(NW - 1) * LD + NW - 1], dor = diag0[NW - 1];
      ok &= dor > 0.0;
      fast_solve &= !(djj > 1e-13 * dor) || djj >= kFastPivot * dor;
      if (tid == NW - 1) invd[NW - 1] = djj > 1e-13 * dor ? fast_rcp(djj) : 0.0;
    }
    bsync();
    return ok;
  }
  // K x = v with K = L D L^T; v in shared memory is overwritten, `tmp` [NW] is scratch.
  // Fast path: x = Linv^T D^-1 Linv v, two products with the inverse of the unit factor (upper triangle of Ks,
  // see factor), TPR threads per row.  The explicit inverse is as accurate as substitution while the factor is
  // well conditioned, and loses what the pivots lost beyond that: measured on the C port (ORC_LINLOG) the two
  // differ by <= 5e-8 relative when every accepted pivot kept >= kFastPivot of its diagonal (94 % of all solves,
  // iteration counts and statuses of 18 000 closed-loop solves identical to the Cholesky reference), and by up to
  // 100 % below that - near convergence of a degenerate node, where the dual residual then stalls.  Those solves
  // take the slow path: substitution on warp 0 with the unscaled columns of the lower triangle (backward stable).
  __device__ __forceinline__ void solve_inplace(double* v, double* tmp) const {
    if (fast_solve) {
      const int row = tid / TPR, part = tid % TPR;
      const bool act = row < NW;
      double s = 0.0;
      if (act) {
#pragma unroll 2
        for (int c = part; c < row; c += TPR) s += Ks[c * LD + row] * v[c];
      }
      s = tpr_sum(s);
      if (act && part == 0) tmp[row] = (v[row] + s) * invd[row];
      bsync();
      s = 0.0;
      if (act) {
#pragma unroll 2
        for (int i = row + 1 + part; i < NW; i += TPR) s += Ks[row * LD + i] * tmp[i];
      }
      s = tpr_sum(s);
      if (act && part == 0) v[row] = tmp[row] + s;
    } else if (wid == 0) {
      const int me = lane < NW ? lane : 0;
      const double myinv = lane < NW ? invd[lane] : 0.0;
      double acc = lane < NW ? v[lane] : 0.0;
#pragma unroll 4
      for (int j = 0; j < NW - 1; ++j) {  // L z = v,  y = D^-1 z:  c_ij y_j = L_ij z_j
        const double yj = __shfl_sync(kFull, acc * myinv, j);
        if (lane > j && lane < NW) acc -= Ks[me * LD + j] * yj;
      }
      const double y = acc * myinv;
      double sum = 0.0;
#pragma unroll 4
      for (int k = NW - 1; k > 0; --k) {  // L' x = y:  x_i = y_i - (1 / d_i) sum_{k > i} c_ki x_k
        const double xk = __shfl_sync(kFull, y - myinv * sum, k);
        if (lane < k) sum += Ks[k * LD + me] * xk;
      }
      if (lane < NW) v[lane] = y - myinv * sum;
    }
    bsync();
  }

  struct QpOut {
    int status, iters;
    double obj, kkt;
  };

  // One row of the complementarity system.  Everything a pass needs from (s, lam, slack, C.dw):
  //   rc = s - slack, ds = -rc - cdw, dl = -lam - (corr + lam ds) / s      (corr = 0 for the predictor)
  struct RowStep {
    double ds, dl, inv_s, inv_l;
  };
  __device__ __forceinline__ static RowStep row_step(double s, double l, double slk, double cdw, double corr) {
    RowStep r;
    const double rsl = fast_rcp(s * l);  // one reciprocal per row and pass
    r.inv_s = l * rsl, r.inv_l = s * rsl;
    r.ds = -(s - slk) - cdw;
    r.dl = -l - (corr + l * r.ds) * r.inv_s;
    return r;
  }

  // block (row, b) of the KKT matrix: acc[c] for c < NZ.
  //   diag:  Hw[ax][rr][c] (first == true) + sum_{q in [q0, q1)} DQ[q] EQ[q][rr] EQ[q][c]
  //   all:   sum_slots M_slot(ax, b) QP[ax][kp][rr] QP[b][kp][c]      (slots == true)
  __device__ __forceinline__ void kkt_block(int ax, int rr, int b, bool first, int q0, int q1, bool slots, double (&acc)[NZ]) const {
#pragma unroll
    for (int c = 0; c < NZ; ++c) acc[c] = first ? T.Hw[ax][rr][c] : 0.0;
#pragma unroll 1
    for (int q = q0; q < q1; ++q) {
      const double f = DQ[ax * NQ + q] * eq_row(ax, q)[rr];
#pragma unroll
      for (int c = 0; c < NZ; ++c) acc[c] += f * eq_row(ax, q)[c];
    }
    if (slots) {
      // symmetric 3x3 block stored as (00,01,02,11,12,22): entry (ax, b)
      const int mi = ax == b ? (ax == 0 ? 0 : ax == 1 ? 3 : 5) : (ax + b == 1 ? 1 : ax + b == 2 ? 2 : 4);
#pragma unroll 1
      for (int slot = 0; slot < nkp; ++slot) {
        const int kp = skp[slot];
        const double f = Mk[6 * slot + mi] * qp_row(ax, kp)[rr];
#pragma unroll
        for (int c = 0; c < NZ; ++c) acc[c] += f * qp_row(b, kp)[c];
      }
    }
  }
  __device__ __forceinline__ static double tpr_sum(double v) {
    if (TPR > 1) {
      v += __shfl_xor_sync(kFull, v, 1);
      v += __shfl_xor_sync(kFull, v, 2);
    }
    return v;
  }

  __device__ QpOut solve_qp() {
    QpOut out{HDSM_MAX_ITER, 0, INFINITY, INFINITY};
    const double tol = A.tol, c0v = s0[9], cutoff = s0[10];
    init_groups();
    init_pairs();
    int nrows = 0;
    for (int idx = tid; idx < NQ3; idx += NT) nrows += sqc[idx] ? 0 : 2;
    if (tid < 2 * nkp) nrows += sege[tid] - segb[tid];
    double cnt4[4] = {(double)nrows, 0, 0, 0};
    reduce4<0>(cnt4);
    const int mtot = (int)cnt4[0];
    const double inv_m = 1.0 / (mtot > 0 ? mtot : 1);
    // KKT row handled by this thread (TPR threads share a row)
    const int row = tid / TPR, part = tid % TPR;
    const bool rowact = row < NW;
    const int ax = rowact ? row / NZ : 0, rr = rowact ? row - ax * NZ : 0;
    // start: unconstrained minimiser of the objective
    if (tid < NW) {
      const int a = tid / NZ, r = tid - a * NZ;
      double v = 0;
#pragma unroll
      for (int c = 0; c < NZ; ++c) v -= T.HwInv[a][r][c] * g[a * NZ + c];
      w[tid] = v;
    }
    double gm4[4] = {tid < NW ? fabs(g[tid]) : 0.0, 0, 0, 0};
    reduce4<1>(gm4);  // includes the barrier that publishes w
    const double gmax = gm4[0];
    positions_of(w, pbar, p);
    quantities_of(w, qbar, qv);
    bsync();
    {
      const int kp = slot_of_thread < nkp ? skp[slot_of_thread] : 0;
      const double px = p[3 * kp], py = p[3 * kp + 1], pz = p[3 * kp + 2];
      for_my_rows([&](const double* r, double& s, double& l) {
        const double slk = r[3] - (r[0] * px + r[1] * py + r[2] * pz);
        s = fmax(slk, 1.0);
        l = 1.0 / s;
      });
    }
#pragma unroll 1
    for (int idx = tid; idx < NQ3; idx += NT) {
      const int a = idx / NQ, q = idx - a * NQ;
      if (sqc[idx]) continue;
      const double sc = 0.05 * (sqhi[idx] - sqlo[idx]);
#pragma unroll
      for (int side = 0; side < 2; ++side) {
        const double s = fmax(box_slack(idx, side, a, q), sc);
        bs[2 * idx + side] = s, bl[2 * idx + side] = 1.0 / s;
      }
    }
    bsync();

    tick(2);
    // The step of one iteration is APPLIED at the top of the next one, in the same sweep over the rows that computes
    // the next residuals: positions and box quantities are linear in w, so p + al dp and q + al dq are what
    // positions_of / quantities_of of the new w would give, and neither a separate update sweep nor the two dense
    // products and their barrier are needed.  (upd: a step is pending; al, smu and the two directions are those of
    // the previous iteration.)
    const int kp = slot_of_thread < nkp ? skp[slot_of_thread] : 0;
    double px = p[3 * kp], py = p[3 * kp + 1], pz = p[3 * kp + 2];
    double dx = 0, dy = 0, dz = 0, ex = 0, ey = 0, ez = 0, al = 0, smu = 0;
    bool upd = false;
#pragma unroll 1
    for (int it = 0;; ++it) {
      tick(8);
      // ---- (pending step, then) pass 1: residuals and barrier blocks
      double acc4[4] = {0, 0, 0, 0};  // mu, rcmax, lam.slack, sum lam
      const double qx = upd ? px + al * ex : px, qy = upd ? py + al * ey : py, qz = upd ? pz + al * ez : pz;  // new positions
      double p_own = 0.0;
      if (upd && tid < K3) p_own = p[tid] + al * dpc[tid];
      {
        double m0 = 0, m1 = 0, m2 = 0, m3 = 0, m4 = 0, m5 = 0, t0 = 0, t1 = 0, t2 = 0, f0 = 0, f1 = 0, f2 = 0;
        for_my_rows([&](const double* r, double& s, double& l) {
          const double nx = r[0], ny = r[1], nz = r[2];
          if (upd) {
            const double slk0 = r[3] - (nx * px + ny * py + nz * pz);
            const double rc0 = s - slk0, inv_s = fast_rcp(s);
            const double dsa = -rc0 - (nx * dx + ny * dy + nz * dz);
            const double dla = -l - l * dsa * inv_s;
            const double ds = -rc0 - (nx * ex + ny * ey + nz * ez);
            const double dl = -l - ((dsa * dla - smu) + l * ds) * inv_s;
            s += al * ds, l += al * dl;
          }
          const double slk = r[3] - (nx * qx + ny * qy + nz * qz);
          const double rc = s - slk, d = l * fast_rcp(s), t = d * rc;
          acc4[0] += s * l, acc4[1] = fmax(acc4[1], fabs(rc)), acc4[2] += l * slk, acc4[3] += l;
          const double dx = d * nx, dy = d * ny, dz = d * nz;
          m0 += dx * nx, m1 += dx * ny, m2 += dx * nz, m3 += dy * ny, m4 += dy * nz, m5 += dz * nz;
          t0 += t * nx, t1 += t * ny, t2 += t * nz;
          f0 += l * nx, f1 += l * ny, f2 += l * nz;
        });
        // twelve independent butterflies per step, inside the groups of lps lanes.  Every lane of the warp takes part
        // (threads without a step carry zeros) so that the mask is the constant full one: with a run-time group mask
        // the compiler guards every shuffle level with a MATCH.ANY / vote sequence
        tick2(11);
        for (int o = lps >> 1; o > 0; o >>= 1) {
          m0 += __shfl_xor_sync(kFull, m0, o), m1 += __shfl_xor_sync(kFull, m1, o), m2 += __shfl_xor_sync(kFull, m2, o);
          m3 += __shfl_xor_sync(kFull, m3, o), m4 += __shfl_xor_sync(kFull, m4, o), m5 += __shfl_xor_sync(kFull, m5, o);
          t0 += __shfl_xor_sync(kFull, t0, o), t1 += __shfl_xor_sync(kFull, t1, o), t2 += __shfl_xor_sync(kFull, t2, o);
          f0 += __shfl_xor_sync(kFull, f0, o), f1 += __shfl_xor_sync(kFull, f1, o), f2 += __shfl_xor_sync(kFull, f2, o);
        }
        if (slot_of_thread < nkp) {
          if (sub == 0) {
            double* M = Mk + 6 * slot_of_thread;
            M[0] = m0, M[1] = m1, M[2] = m2, M[3] = m3, M[4] = m4, M[5] = m5;
            Tk[3 * slot_of_thread] = t0, Tk[3 * slot_of_thread + 1] = t1, Tk[3 * slot_of_thread + 2] = t2;
            Fk[3 * slot_of_thread] = f0, Fk[3 * slot_of_thread + 1] = f1, Fk[3 * slot_of_thread + 2] = f2;
          }
        }
        tick2(12);
      }
#pragma unroll 1
      for (int idx = tid; idx < NQ3; idx += NT) {
        double dsum = 0, tsum = 0, fsum = 0;
        if (!sqc[idx]) {
          const double irange = fast_rcp(sqhi[idx] - sqlo[idx]);
          const double qold = qv[idx], qnew = upd ? qold + al * dqc[idx] : qold, dqa = upd ? dq[idx] : 0.0, dqb = upd ? dqc[idx] : 0.0;
#pragma unroll
          for (int side = 0; side < 2; ++side) {
            const double sg = side == 0 ? 1.0 : -1.0;
            double s = bs[2 * idx + side], l = bl[2 * idx + side];
            if (upd) {
              const double slk0 = side == 0 ? sqhi[idx] - qold : qold - sqlo[idx], inv_s = fast_rcp(s);
              const double rc0 = s - slk0, dsa = -rc0 - sg * dqa, dla = -l - l * dsa * inv_s;
              const double ds = -rc0 - sg * dqb;
              const double dl = -l - ((dsa * dla - smu) + l * ds) * inv_s;
              s += al * ds, l += al * dl;
              bs[2 * idx + side] = s, bl[2 * idx + side] = l;
            }
            const double slk = side == 0 ? sqhi[idx] - qnew : qnew - sqlo[idx];
            const double rc = s - slk, d = l * fast_rcp(s);
            acc4[0] += s * l, acc4[1] = fmax(acc4[1], fabs(rc) * irange), acc4[2] += l * slk, acc4[3] += l;
            dsum += d, tsum += sg * d * rc, fsum += sg * l;
          }
          if (upd) qv[idx] = qnew;  // read and written by this thread only
        }
        DQ[idx] = dsum, TQ[idx] = tsum, FQ[idx] = fsum;
      }
      if (upd && tid < NW) {  // nobody reads w or dwc in this phase
        const double v = w[tid] + al * dwc[tid];
        if (!isfinite(v)) acc4[1] = INFINITY;
        w[tid] = v;
      }
      px = qx, py = qy, pz = qz;
      __syncwarp();
      tick2(13);
      reduce4<2>(acc4);  // also the barrier that publishes Mk / Tk / Fk / DQ / TQ / FQ and w
      if (upd && tid < K3) p[tid] = p_own;  // for publish(): every reader of p in here has it in registers, and the next barrier comes before any return
      if (!(acc4[1] < INFINITY)) {  // a step produced a non-finite w
        out.status = HDSM_NUMERICAL, out.iters = it > 0 ? it - 1 : 0;
        return out;
      }
      const double mu = acc4[0] * inv_m, rcmax = acc4[1], lamsl = acc4[2], lamsum = acc4[3];

      tick(3);
      // ---- smooth gradient, dual residual, objective, right-hand side of the predictor
      double hg = 0, fi = 0, ti = 0;
      if (rowact) {
#pragma unroll 1
        for (int c = part; c < NZ; c += TPR) hg += T.Hw[ax][rr][c] * w[ax * NZ + c];
#pragma unroll 1
        for (int slot = part; slot < nkp; slot += TPR) {
          const double qp = qp_row(ax, skp[slot])[rr];
          fi += Fk[3 * slot + ax] * qp, ti += Tk[3 * slot + ax] * qp;
        }
#pragma unroll 2
        for (int q = part; q < NQ; q += TPR) {
          const double e = eq_row(ax, q)[rr];
          fi += FQ[ax * NQ + q] * e, ti += TQ[ax * NQ + q] * e;
        }
      }
      hg = tpr_sum(hg), fi = tpr_sum(fi), ti = tpr_sum(ti);
      if (rowact) hg += g[row];
      const bool owner = rowact && part == 0;
      if (owner && cutoff < INFINITY) dwc[row] = g[row] + fi;  // v = g + C'lam for the dual bound below (dwc is free here)
      const double wi = owner ? w[row] : 0.0;
      double r4[4] = {owner ? fabs(hg + fi) : 0.0, owner ? fabs(fi) : 0.0, wi * fi, owner ? wi * (hg + g[row]) : 0.0};
      reduce4<3>(r4);
      const double rdmax = r4[0], fmaxv = r4[1], wf = r4[2];
      const double obj = c0v + 0.5 * r4[3];
      const auto accept = [&](double tl) {
        return rdmax <= tl * (1 + gmax) && rcmax <= tl && mu <= 0.1 * tl * fmax(1.0, fabs(obj));
      };
      if (accept(tol) || (it >= A.max_iter && accept(kLooseTol))) {
        out.status = HDSM_OPTIMAL, out.iters = it, out.obj = obj;
        out.kkt = fmax(rdmax / (1 + gmax), fmax(rcmax, mu / fmax(1.0, fabs(obj))));  // scaled as in SURVEY 8(d)
        return out;
      }
      // Lagrangian dual bound (weak duality, any lam >= 0): min_w L(w, lam) = c0 - 1/2 v'Hw^-1 v - d'lam with
      // v = g + C'lam (published in dwc by the reduction above) and d'lam = lam'(d - Cw) + w'C'lam.  Once it
      // reaches the incumbent this relaxation cannot improve it: the solve is abandoned.
      if (cutoff < INFINITY) {
        double q4[4] = {0, 0, 0, 0};
        if (owner) {
          double hv = 0;
#pragma unroll
          for (int c = 0; c < NZ; ++c) hv += T.HwInv[ax][rr][c] * dwc[ax * NZ + c];
          q4[0] = dwc[row] * hv;
        }
        reduce4<0>(q4);
        if (c0v - 0.5 * q4[0] - (lamsl + wf) >= cutoff) {
          out.status = kCutoff, out.iters = it;
          return out;
        }
      }
      if (mtot > 0 && lamsum > 0 && it >= 3) {  // Farkas certificate: lam >= 0, C'lam ~ 0, d'lam < 0
        const double dlam = (lamsl + wf) / lamsum;
        if (fmaxv / lamsum < 1e-9 * fmax(1.0, -dlam * 1e3) && dlam < -1e-7) {
          out.status = HDSM_INFEASIBLE, out.iters = it;
          return out;
        }
      }
      if (it >= A.max_iter) {
        out.iters = it;
        return out;
      }

      tick(4);
      // ---- K = Hw + sum_slots QP' M QP + sum_q DQ EQ EQ'   (lower blocks, written to shared memory)
      if (TPR == 1) {
        if (rowact) {
#pragma unroll 1
          for (int b = 0; b < 3; ++b) {
            double acc[NZ];
            kkt_block(ax, rr, b, b == ax, 0, b == ax ? NQ : 0, true, acc);
#pragma unroll
            for (int c = 0; c < NZ; ++c) Ks[row * LD + b * NZ + c] = acc[c];
            if (b == ax) {
              double dg = acc[0];
#pragma unroll
              for (int c = 1; c < NZ; ++c) dg = c == rr ? acc[c] : dg;
              diag0[row] = dg;
            }
          }
        }
      } else {
        // four threads per row: part b < ax -> off-diagonal block b; part ax -> diagonal block with the
        // first half of the box terms and the position terms; part 3 -> second half of the box terms
        double acc[NZ];
        const bool diag_main = rowact && part == ax, diag_aux = rowact && part == 3;
        const bool off = rowact && part < ax;
        if (diag_main) kkt_block(ax, rr, ax, true, 0, NQ / 2, true, acc);
        else if (diag_aux) kkt_block(ax, rr, ax, false, NQ / 2, NQ, false, acc);
        else if (off) kkt_block(ax, rr, part, false, 0, 0, true, acc);
        else {
#pragma unroll
          for (int c = 0; c < NZ; ++c) acc[c] = 0.0;
        }
#pragma unroll
        for (int c = 0; c < NZ; ++c) {
          const double o = __shfl_sync(kFull, acc[c], (lane & ~3) | 3);
          if (diag_main) acc[c] += o;
        }
        if (diag_main || off) {
          const int b = diag_main ? ax : part;
#pragma unroll
          for (int c = 0; c < NZ; ++c) Ks[row * LD + b * NZ + c] = acc[c];
        }
        if (diag_main) {
          double dg = acc[0];
#pragma unroll
          for (int c = 1; c < NZ; ++c) dg = c == rr ? acc[c] : dg;
          diag0[row] = dg;
        }
      }
      if (owner) dw[row] = -hg - ti;  // right-hand side of the predictor
      bsync();
      tick(5);
      if (!factor()) {
        out.status = HDSM_NUMERICAL, out.iters = it;
        return out;
      }

      tick(6);
      // ---- predictor
      solve_inplace(dw, dwc);  // dwc is free until the corrector's right-hand side is written
      tick(7);
      positions_of(dw, nullptr, dp);
      quantities_of(dw, nullptr, dq);
      bsync();
      dx = dp[3 * kp], dy = dp[3 * kp + 1], dz = dp[3 * kp + 2];
      tick(8);
      // largest relative decrease of any s or lam along the step: alpha_max = 1 / rmax
      double a4[4] = {0, 0, 0, 0};  // rmax, sum s.lam, sum (s dl + l ds), sum ds.dl
      for_my_rows([&](const double* r, double& s, double& l) {
        const double slk = r[3] - (r[0] * px + r[1] * py + r[2] * pz);
        const RowStep e = row_step(s, l, slk, r[0] * dx + r[1] * dy + r[2] * dz, 0.0);
        a4[0] = fmax(a4[0], fmax(-e.ds * e.inv_s, -e.dl * e.inv_l));
        a4[1] += s * l, a4[2] += s * e.dl + l * e.ds, a4[3] += e.ds * e.dl;
      });
#pragma unroll 1
      for (int idx = tid; idx < NQ3; idx += NT) {
        const int a = idx / NQ, q = idx - a * NQ;
        if (sqc[idx]) continue;
#pragma unroll
        for (int side = 0; side < 2; ++side) {
          const double s = bs[2 * idx + side], l = bl[2 * idx + side], slk = box_slack(idx, side, a, q);
          const RowStep e = row_step(s, l, slk, side == 0 ? dq[idx] : -dq[idx], 0.0);
          a4[0] = fmax(a4[0], fmax(-e.ds * e.inv_s, -e.dl * e.inv_l));
          a4[1] += s * l, a4[2] += s * e.dl + l * e.ds, a4[3] += e.ds * e.dl;
        }
      }
      __syncwarp();
      reduce4<1>(a4);
      tick2(14);
      const double alpha = a4[0] > 1.0 ? 1.0 / a4[0] : 1.0;
      const double mu_aff = fmax((a4[1] + alpha * a4[2] + alpha * alpha * a4[3]) * inv_m, 0.0);
      const double ratio = mu > 0 ? mu_aff / mu : 0.0;
      smu = ratio * ratio * ratio * mu;

      // ---- corrector right-hand side
      {
        double t0 = 0, t1 = 0, t2 = 0;
        for_my_rows([&](const double* r, double& s, double& l) {
          const double slk = r[3] - (r[0] * px + r[1] * py + r[2] * pz);
          const RowStep e = row_step(s, l, slk, r[0] * dx + r[1] * dy + r[2] * dz, 0.0);
          const double t = (l * (s - slk) - (e.ds * e.dl - smu)) * e.inv_s;
          t0 += t * r[0], t1 += t * r[1], t2 += t * r[2];
        });
        for (int o = lps >> 1; o > 0; o >>= 1)  // all lanes, full mask (see pass 1)
          t0 += __shfl_xor_sync(kFull, t0, o), t1 += __shfl_xor_sync(kFull, t1, o), t2 += __shfl_xor_sync(kFull, t2, o);
        if (slot_of_thread < nkp && sub == 0)
          Tk[3 * slot_of_thread] = t0, Tk[3 * slot_of_thread + 1] = t1, Tk[3 * slot_of_thread + 2] = t2;
      }
#pragma unroll 1
      for (int idx = tid; idx < NQ3; idx += NT) {
        const int a = idx / NQ, q = idx - a * NQ;
        double tsum = 0;
        if (!sqc[idx]) {
#pragma unroll
          for (int side = 0; side < 2; ++side) {
            const double s = bs[2 * idx + side], l = bl[2 * idx + side], slk = box_slack(idx, side, a, q);
            const RowStep e = row_step(s, l, slk, side == 0 ? dq[idx] : -dq[idx], 0.0);
            const double t = (l * (s - slk) - (e.ds * e.dl - smu)) * e.inv_s;
            tsum += side == 0 ? t : -t;
          }
        }
        TQ[idx] = tsum;
      }
      bsync();
      tick2(15);
      double tc = 0;
      if (rowact) {
#pragma unroll 1
        for (int slot = part; slot < nkp; slot += TPR) tc += Tk[3 * slot + ax] * qp_row(ax, skp[slot])[rr];
#pragma unroll 2
        for (int q = part; q < NQ; q += TPR) tc += TQ[ax * NQ + q] * eq_row(ax, q)[rr];
      }
      tc = tpr_sum(tc);
      if (owner) dwc[row] = -hg - tc;
      bsync();
      tick(9);
      solve_inplace(dwc, dw);  // the predictor direction lives on in dp / dq only
      tick(7);
      positions_of(dwc, nullptr, dpc);
      quantities_of(dwc, nullptr, dqc);
      bsync();
      ex = dpc[3 * kp], ey = dpc[3 * kp + 1], ez = dpc[3 * kp + 2];

      tick(8);
      // ---- step length of the combined direction
      double m4[4] = {0, 0, 0, 0};
      for_my_rows([&](const double* r, double& s, double& l) {
        const double slk = r[3] - (r[0] * px + r[1] * py + r[2] * pz);
        const double dsa = -(s - slk) - (r[0] * dx + r[1] * dy + r[2] * dz);
        const RowStep e0 = row_step(s, l, slk, r[0] * ex + r[1] * ey + r[2] * ez, 0.0);
        const double dla = -l - l * dsa * e0.inv_s;
        const double dl = -l - ((dsa * dla - smu) + l * e0.ds) * e0.inv_s;
        m4[0] = fmax(m4[0], fmax(-e0.ds * e0.inv_s, -dl * e0.inv_l));
      });
#pragma unroll 1
      for (int idx = tid; idx < NQ3; idx += NT) {
        const int a = idx / NQ, q = idx - a * NQ;
        if (sqc[idx]) continue;
#pragma unroll
        for (int side = 0; side < 2; ++side) {
          const double s = bs[2 * idx + side], l = bl[2 * idx + side], slk = box_slack(idx, side, a, q);
          const double sg = side == 0 ? 1.0 : -1.0;
          const double dsa = -(s - slk) - sg * dq[idx];
          const RowStep e0 = row_step(s, l, slk, sg * dqc[idx], 0.0);
          const double dla = -l - l * dsa * e0.inv_s;
          const double dl = -l - ((dsa * dla - smu) + l * e0.ds) * e0.inv_s;
          m4[0] = fmax(m4[0], fmax(-e0.ds * e0.inv_s, -dl * e0.inv_l));
        }
      }
      __syncwarp();
      reduce4<1>(m4);
      tick2(14);
      // 0.97 of the way to the boundary: 0.995 leaves the blocking pair so far off the central path
      // that predictor and centring steps alternate without reducing mu on ~0.4% of the QPs
      al = m4[0] > kStepFrac ? kStepFrac / m4[0] : 1.0;
      upd = true;  // applied by the first sweep of the next iteration
      tick(9);
    }
  }

  // ---------------------------------------------------------------- K3 + epilogue
  // The search runs in ROUNDS of up to A.width open nodes (width 1: plain depth-first search).  The nodes of a
  // round are solved independently, all against the incumbent the round started with, and merged in a fixed
  // order - incumbents in pop order, then children pushed so that the first node's children end up on top of
  // the stack - so the result depends on the width only, never on who solved what or when.  That leaves the
  // execution free: a round's nodes are dealt to the `csize` thread blocks of the agent's cluster (node j to
  // block j mod csize); every block keeps an identical copy of the stack and the incumbent, solves its nodes,
  // publishes each outcome (status, objective, branching step and children, or the integral solution) in its
  // own shared memory, and after one cluster barrier reads the others' through distributed shared memory.
  // csize = 1 runs the same rounds sequentially in one block.  Warp 0 owns the search state; the other warps
  // only join solve_qp.  ctl[0]: 0 = finished, 1 = a node is ready to be solved, 2 = nothing for this block.
  static constexpr int kMsgDoubles = 6 + (N + 1) / 2 + NW;  // header (2), obj, kkt, children (1), pad, fullsig, w
  struct Msg {  // view of one outcome slot
    double* d;
    __device__ int& status() const { return reinterpret_cast<int*>(d)[0]; }
    __device__ int& iters() const { return reinterpret_cast<int*>(d)[1]; }
    __device__ int& bk() const { return reinterpret_cast<int*>(d)[2]; }      // branching step, -1: integral
    __device__ int& nchild() const { return reinterpret_cast<int*>(d)[3]; }
    __device__ double& obj() const { return d[2]; }
    __device__ double& kkt() const { return d[3]; }
    __device__ unsigned char* child() const { return reinterpret_cast<unsigned char*>(d + 4); }  // masks of step bk, push order
    __device__ int& rows() const { return reinterpret_cast<int*>(d + 5)[0]; }
    __device__ int* sig() const { return reinterpret_cast<int*>(d + 6); }
    __device__ double* w() const { return d + 6 + (N + 1) / 2; }
  };
  double* xch;  // [2 parities][kMaxWidth] slots of kMsgDoubles, behind the row pool

  __device__ __forceinline__ static void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  // generic pointer to `mine` in the shared memory of block `rank` of this cluster
  __device__ __forceinline__ static double* peer(double* mine, unsigned rank) {
    unsigned long long out;
    asm volatile("mapa.u64 %0, %1, %2;" : "=l"(out) : "l"(reinterpret_cast<unsigned long long>(mine)), "r"(rank));
    return reinterpret_cast<double*>(out);
  }

  // warp 0, after a node's relaxation was solved: coverage of every segment (p_k, p_k+1) by one member of its
  // candidate set; integral -> the solution goes into the slot, otherwise the step to branch on and the children
  __device__ void publish(const QpOut& q, Msg m, int nrows) {
    if (lane == 0) m.status() = q.status, m.iters() = q.iters, m.obj() = q.obj, m.kkt() = q.kkt, m.bk() = -1, m.nchild() = 0, m.rows() = nrows;
    __syncwarp();
    if (q.status != HDSM_OPTIMAL) return;
    for (int idx = lane; idx < N * Peff; idx += 32) {
      const int k = idx / Peff, j = idx - k * Peff;
      double v = INFINITY;
      if (cur[k] >> j & 1) {
        v = -INFINITY;
        const double ax_ = p[3 * k], ay = p[3 * k + 1], az = p[3 * k + 2];
        const double bx = p[3 * k + 3], by = p[3 * k + 4], bz = p[3 * k + 5];
        for (int r = 0; r < prow_n[j]; ++r) {
          const double* c = poly + 4 * (j * A.rmax + r);
          v = fmax(v, fmax(c[0] * ax_ + c[1] * ay + c[2] * az, c[0] * bx + c[1] * by + c[2] * bz) - c[3]);
        }
      }
      viol[k * kMaxP + j] = v;
    }
    __syncwarp();
    // branch on the uncovered step that is farthest from all of its candidates
    int bk = -1;
    double bkv = 0.0;
#pragma unroll 1
    for (int k = 0; k < N; ++k) {
      int fk = -1;
      double vmin = INFINITY;
#pragma unroll 1
      for (int j = 0; j < Peff; ++j) {
        const double v = viol[k * kMaxP + j];
        vmin = fmin(vmin, v);
        if (fk < 0 && v <= kContainTol) fk = j;
      }
      if (lane == 0) m.sig()[k] = fk;
      if (fk < 0 && (bk < 0 || vmin > bkv)) bk = k, bkv = vmin;
    }
    if (bk < 0) {  // node optimum is feasible for the mixed-integer problem
      if (lane < NW) m.w()[lane] = w[lane];
      __syncwarp();
      return;
    }
    // split the candidate set of step bk in two halves ordered by (violation, index), least violated explored
    // first.  A half whose hull still contains the segment of the node optimum would be solved to the very same
    // point and then split again on the same step: that solve is skipped, the half is split right away.
    int order[kMaxP], no = 0;
    double vv[kMaxP];
    for (int j = 0; j < Peff; ++j)
      if (cur[bk] >> j & 1) order[no] = j, vv[no++] = viol[bk * kMaxP + j];
    for (int x = 1; x < no; ++x)
      for (int y = x; y > 0 && vv[y] < vv[y - 1]; --y) {
        const double tv = vv[y];
        vv[y] = vv[y - 1], vv[y - 1] = tv;
        const int to = order[y];
        order[y] = order[y - 1], order[y - 1] = to;
      }
    int nchild = 0;
    if (no > 1) {
      int wx0[2 * kMaxP], wx1[2 * kMaxP], nw = 0;
      const int h = (no + 1) / 2;
      wx0[nw] = 0, wx1[nw++] = h;  // last in, first out: the upper half is pushed on the search stack first
      wx0[nw] = h, wx1[nw++] = no;
      const double sax = p[3 * bk], say = p[3 * bk + 1], saz = p[3 * bk + 2];
      const double sbx = p[3 * bk + 3], sby = p[3 * bk + 4], sbz = p[3 * bk + 5];
#pragma unroll 1
      while (nw > 0) {
        const int x0 = wx0[--nw], x1 = wx1[nw];
        unsigned mm = 0;
        for (int x = x0; x < x1; ++x) mm |= 1u << order[x];
        bool contains = false;
        if (x1 - x0 > 1) {
          bool out = false;
          for (int i = lane; i < A.P * A.rmax; i += 32) {
            double n[3], b;
            if (hull_row(mm, i, n, b))
              out |= fmax(n[0] * sax + n[1] * say + n[2] * saz, n[0] * sbx + n[1] * sby + n[2] * sbz) - b > kContainTol;
          }
          contains = !__any_sync(kFull, out);
        }
        if (contains) {
          const int hh = (x1 - x0 + 1) / 2;
          wx0[nw] = x0, wx1[nw++] = x0 + hh;
          wx0[nw] = x0 + hh, wx1[nw++] = x1;
          continue;
        }
        if (lane == 0) m.child()[nchild] = (unsigned char)mm;
        ++nchild;
      }
    }
    if (lane == 0) m.bk() = bk, m.nchild() = nchild;
    __syncwarp();
  }

  __device__ void run(int agent, int crank, int csize) {
    hdsm_result R{HDSM_INFEASIBLE, 0, 0, 0, INFINITY, INFINITY};
    int st = -1, top = 0, nodes = 0, iters = 0, maxrows = 0, fail = 0, parity = 0, rounds = 0;
    bool deferred = false;
    double best = INFINITY, bestkkt = INFINITY;
    bool exhausted = true, overflow = false;
    const int width = min(max(A.width, 1), kMaxWidth);
    xch = rl + A.row_cap + (A.P * A.rmax + 7) / 8;
    unsigned char* pop = reinterpret_cast<unsigned char*>(xch + 2 * kMaxWidth * kMsgDoubles);  // [kMaxWidth][16] masks of the round
    tick_init();
    load_and_index(agent);
    if (wid == 0) {
      tick1(11);
      st = setup(agent);
      if (lane == 0) ctl[1] = st;
      tick1(12);
    }
    bsync();
    if (ctl[1] < 0) {  // uniform over the block
      scan_neighbours(agent);
      tick1(13);
      const int sn = build_neighbour_rows(agent);
      if (wid == 0) st = sn;
    }
    if (wid == 0) {
      tick1(14);
      if (st < 0 && !root_sets(agent)) st = HDSM_INFEASIBLE;
      if (lane == 0) ctl[1] = st;
      tick1(15);
    }
    bsync();
    if (ctl[1] < 0 && !(A.dbg & 1)) dominance_filter();  // uniform over the block
    if (wid == 0) {
      if (st < 0) {
        if (lane < 16) stack[lane] = lane < N ? cur[lane] : 0;
        if (lane == 0) sbnd[0] = -INFINITY;
        top = 1;
        __syncwarp();
      }
      tick(0);
    }
    int cnt = 0;       // nodes of the current round (warp 0)
    double root_obj = INFINITY;
    for (;;) {
      // ---- pop the round (identical in every block of the cluster)
      if (wid == 0) {
        cnt = 0;
        if (st < 0 && !overflow && A.round_budget > 0 && rounds >= A.round_budget && top > 0) deferred = true;
        if (st < 0 && !overflow && !deferred) {
          while (exhausted && cnt == 0 && top > 0) {
            while (cnt < width && top > 0) {
              if (nodes + cnt >= A.max_nodes) {
                if (cnt == 0) exhausted = false;
                break;
              }
              --top;
              // a node whose parent's optimum already reaches the incumbent cannot improve it: dropped unsolved
              if (!(A.dbg & 2) && sbnd[top] >= best - kPruneRel * fmax(1.0, fabs(best))) continue;
              if (lane < 16) pop[cnt * 16 + lane] = stack[top * 16 + lane];
              ++cnt;
            }
          }
          __syncwarp();
        }
        if (lane == 0) {
          ctl[0] = cnt > 0 ? 1 : 0;
          ctl[2] = cnt;
          s0[10] = best < INFINITY ? best - kPruneRel * fmax(1.0, fabs(best)) : INFINITY;  // cutoff of the dual bound
        }
      }
      bsync();
      if (ctl[0] == 0) break;
      const int rcnt = ctl[2];
      // ---- solve this block's nodes of the round, one after the other
#pragma unroll 1
      for (int j = crank; j < rcnt; j += csize) {
        Msg m{xch + (parity * kMaxWidth + j) * kMsgDoubles};
        if (wid == 0) {
          if (lane < 16) cur[lane] = pop[j * 16 + lane];
          __syncwarp();
          tick(10);
          const int nstat = build_static_rows();
          tick(1);
          if (lane == 0) ctl[0] = nstat < 0 ? 2 : 1;
          if (nstat < 0 && lane == 0)  // the row pool is too small for this agent: the large-memory pass redoes it
            m.status() = HDSM_ROW_OVERFLOW, m.iters() = 0, m.bk() = -1, m.nchild() = 0, m.rows() = 0, m.obj() = INFINITY;
          if (nstat >= 0) maxrows = max(maxrows, nstat + n_nbr_rows);
        }
        bsync();
        if (ctl[0] == 1) {  // uniform over the block
          tick_start();
          const QpOut q = solve_qp();  // all warps; the result is uniform over the block
          if (wid == 0) publish(q, m, maxrows);
        }
        bsync();
      }
      if (csize > 1) cluster_sync();  // outcomes of all blocks visible (release / acquire at cluster scope)
      if (wid != 0) continue;
      // ---- merge 1: incumbents, in pop order
#pragma unroll 1
      for (int j = 0; j < cnt; ++j) {
        double* slot = xch + (parity * kMaxWidth + j) * kMsgDoubles;
        Msg m{csize > 1 ? peer(slot, j % csize) : slot};
        const int ms = m.status();
        ++nodes;
        iters += m.iters();
        maxrows = max(maxrows, m.rows());
        if (ms == HDSM_ROW_OVERFLOW) overflow = true;
        if (ms != HDSM_OPTIMAL) {
          if (ms != HDSM_INFEASIBLE && ms != kCutoff && ms != HDSM_ROW_OVERFLOW) fail = ms;
          continue;
        }
        const double mo = m.obj();
        if (rounds == 0 && j == 0) root_obj = mo;
        if (m.bk() >= 0 || mo >= best - kPruneRel * fmax(1.0, fabs(best))) continue;
        best = mo, bestkkt = m.kkt();
        if (lane < NW) bestw[lane] = m.w()[lane];
        if (lane < N) bestsig[lane] = m.sig()[lane];
        __syncwarp();
      }
      // ---- merge 2: children, last node first so that the first node's children end up on top of the stack
#pragma unroll 1
      for (int j = cnt - 1; j >= 0 && exhausted && !overflow; --j) {
        double* slot = xch + (parity * kMaxWidth + j) * kMsgDoubles;
        Msg m{csize > 1 ? peer(slot, j % csize) : slot};
        if (m.status() != HDSM_OPTIMAL || m.bk() < 0) continue;
        const double mo = m.obj();
        if (mo >= best - kPruneRel * fmax(1.0, fabs(best))) continue;
        const int bk = m.bk(), nc = m.nchild();
#pragma unroll 1
        for (int i = 0; i < nc; ++i) {
          if (top + 1 > kStackCap) {  // stack full: the optimum stays unproven
            exhausted = false;
            break;
          }
          const unsigned char cm = m.child()[i];
          if (lane < 16) stack[top * 16 + lane] = lane == bk ? cm : pop[j * 16 + lane];
          if (lane == 0) sbnd[top] = mo;
          ++top;
        }
        __syncwarp();
      }
      // ---- warm start (optional): the root has branched - the previous plan shifted by one step names a cell per step
      // (the member of the root's candidate set the segment (prev[k+1], prev[k+2]) lies deepest in, if it lies in one);
      // that assignment goes on top of the stack, the next round solves it first and an incumbent exists early
      if (A.warm_start && rounds == 0 && top > 0 && exhausted && !overflow && top + 1 <= kStackCap) {
        const double* prev = A.prev + (size_t)agent * K3;
        for (int idx = lane; idx < N * Peff; idx += 32) {
          const int k = idx / Peff, j = idx - k * Peff;
          double v = INFINITY;
          if (pop[k] >> j & 1) {
            v = -INFINITY;
            const double* a0 = prev + 3 * min(k + 1, N);
            const double* a1 = prev + 3 * min(k + 2, N);
            for (int r = 0; r < prow_n[j]; ++r) {
              const double* c = poly + 4 * (j * A.rmax + r);
              v = fmax(v, fmax(c[0] * a0[0] + c[1] * a0[1] + c[2] * a0[2], c[0] * a1[0] + c[1] * a1[1] + c[2] * a1[2]) - c[3]);
            }
          }
          viol[k * kMaxP + j] = v;
        }
        __syncwarp();
        unsigned hm = 0;
        if (lane < N) {
          int bj = -1;
          double bv = INFINITY;
          for (int j = 0; j < Peff; ++j) {
            const double v = viol[lane * kMaxP + j];
            if (v < bv) bv = v, bj = j;
          }
          if (bj >= 0 && bv <= kContainTol) hm = 1u << bj;
        }
        const bool okh = __all_sync(kFull, lane >= N || hm != 0);
        if (okh) {
          if (lane < 16) stack[top * 16 + lane] = (unsigned char)hm;
          if (lane == 0) sbnd[top] = root_obj;
          ++top;
        }
        __syncwarp();
      }
      parity ^= 1;
      ++rounds;
    }
    if (csize > 1) cluster_sync();  // no block leaves while another may still read its outcome slots
    if (wid != 0 || crank != 0) return;
    if (st < 0) {
      R.nodes = nodes, R.iters = iters, R.rows = maxrows;
      if (overflow) {
        best = INFINITY;
        R.status = A.overflow_status;
      } else if (deferred) {  // the cluster pass of this call starts this agent over (same rounds, same result)
        best = INFINITY;
        R.status = kDeferred;
      } else if (best < INFINITY) {
        R.status = (exhausted && !fail) ? HDSM_OPTIMAL : HDSM_NODE_LIMIT;  // a lost node leaves the optimum unproven
        R.obj = best, R.kkt_res = bestkkt;
      } else {
        R.status = !exhausted ? HDSM_NODE_LIMIT : (fail ? fail : HDSM_INFEASIBLE);
      }
    } else {
      R.status = st;
    }
    tick(10);
    tick_flush(agent);
    write_outputs(agent, R, best < INFINITY);
  }

  // read-back (agent_class.cpp:962-987) and K4 position pack; failed agents get the shifted previous
  // plan in pos_out (:1004-1013) and zeros elsewhere
  __device__ void write_outputs(int agent, const hdsm_result& R, bool have) {
    double* traj = A.traj + (size_t)agent * (N + 1) * 9;
    double* ctrl = A.ctrl + (size_t)agent * N * 3;
    for (int idx = lane; idx < K3; idx += 32) {
      const int k = idx / 3, a = idx - 3 * k;
      double pp = 0, vv = 0, aa = 0;
      if (have) {
        pp = pbar[idx];
        vv = T.cV[a][k][0] * s0[a] + T.cV[a][k][1] * s0[3 + a] + T.cV[a][k][2] * s0[6 + a];
        aa = T.cA[a][k][0] * s0[a] + T.cA[a][k][1] * s0[3 + a] + T.cA[a][k][2] * s0[6 + a];
#pragma unroll
        for (int z = 0; z < NZ; ++z) {
          const double wz = bestw[a * NZ + z];
          pp += T.QP[a][k][z] * wz, vv += T.QV[a][k][z] * wz, aa += T.QA[a][k][z] * wz;
        }
        if (k == 0) pp = s0[a], vv = s0[3 + a], aa = s0[6 + a];  // fixed to state_curr_ (:886-889)
        if (k == N) vv = 0.0, aa = 0.0;                          // fixed to zero (:2078-2081)
      }
      traj[k * 9 + a] = pp, traj[k * 9 + 3 + a] = vv, traj[k * 9 + 6 + a] = aa;
      if (A.pos_out) {
        const double* prev = A.prev + (size_t)agent * K3;
        A.pos_out[(size_t)agent * K3 + idx] = have ? pp : prev[3 * (k < N ? k + 1 : N) + a];
      }
    }
    for (int idx = lane; idx < 3 * N; idx += 32) {
      const int k = idx / 3, a = idx - 3 * k;
      double u = 0;
      if (have) {
        u = T.Up[a][k][0] * s0[a] + T.Up[a][k][1] * s0[3 + a] + T.Up[a][k][2] * s0[6 + a];
#pragma unroll
        for (int z = 0; z < NZ; ++z) u += T.Z[a][k][z] * bestw[a * NZ + z];
      }
      ctrl[idx] = u;
    }
    if (lane < N) A.assign_out[(size_t)agent * N + lane] = have ? bestsig[lane] : -1;
    if (lane < A.P) {
      int used = 0;
      if (have)
        for (int k = 0; k < N; ++k) used |= bestsig[k] == lane;
      A.poly_used[(size_t)agent * A.P + lane] = (uint8_t)used;
    }
    if (lane == 0) A.res[agent] = R;
  }
};

template <int N, int W>
__global__ void __launch_bounds__(32 * W, W == 4 ? HDSM_MINBLOCKS : 1) hdsm_solve_kernel(const Tables* __restrict__ tables, const KernelArgs args) {
  extern __shared__ double smem[];
  // args.csize consecutive blocks (one thread-block cluster when csize > 1) work on one agent
  const int lslot = (int)blockIdx.x / args.csize, crank = (int)blockIdx.x - lslot * args.csize, slot = lslot + args.slot0;
  if (slot >= args.n_local) return;
  const int agent = args.order ? args.order[slot] : slot;
  if (args.only_mask && !((args.only_mask >> (args.res[agent].status & 31)) & 1u)) return;
  Solver<N, W> s(*tables, args, smem);
  s.run(agent, crank, args.csize);
}

// Dispatch order for the next call on the same slots: agents sorted by the interior-point iterations they
// needed this time, most expensive first (counting sort, one block).  A few agents need 100x the median
// work (deep branch and bound); started last they leave the GPU almost idle while they finish, started
// first they overlap with everybody else.  Hardness persists from one replanning step to the next, so last
// step's count is the predictor.  Only the order of execution changes, never a result.
__global__ void __launch_bounds__(1024) hdsm_order_kernel(const hdsm_result* __restrict__ res, int n, int32_t* __restrict__ order) {
  __shared__ int hist[1024];
  __shared__ int wsum[32];
  const int tid = threadIdx.x;
  hist[tid] = 0;
  __syncthreads();
  for (int i = tid; i < n; i += 1024) atomicAdd(&hist[1023 - min(max(res[i].iters, 0), 1023)], 1);
  __syncthreads();
  const int v = hist[tid];  // exclusive prefix sum over the 1024 bins
  int incl = v;
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(kFull, incl, o);
    if ((tid & 31) >= o) incl += t;
  }
  if ((tid & 31) == 31) wsum[tid >> 5] = incl;
  __syncthreads();
  if (tid < 32) {
    int w = wsum[tid];
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(kFull, w, o);
      if (tid >= o) w += t;
    }
    wsum[tid] = w;
  }
  __syncthreads();
  hist[tid] = incl - v + (tid >= 32 ? wsum[(tid >> 5) - 1] : 0);
  __syncthreads();
  for (int i = tid; i < n; i += 1024) order[atomicAdd(&hist[1023 - min(max(res[i].iters, 0), 1023)], 1)] = i;
}

// Centre and radius of every agent's plan points 1..N (the points the inter-agent planes are built from): one
// thread per agent, run once per call in front of the solver kernel, whose neighbour scan then rejects far
// candidates on these 32 bytes.  The radius is rounded up; agents without a plan get radius -1.
__global__ void hdsm_plan_bounds_kernel(int n_rob, int N, const double* __restrict__ all_pos, const uint8_t* __restrict__ all_valid,
                                        double* __restrict__ bounds) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_rob) return;
  const double* q = all_pos + (size_t)j * 3 * (N + 1);
  double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (int k = 1; k <= N; ++k)
    for (int a = 0; a < 3; ++a) lo[a] = fmin(lo[a], q[3 * k + a]), hi[a] = fmax(hi[a], q[3 * k + a]);
  const double c[3] = {0.5 * (lo[0] + hi[0]), 0.5 * (lo[1] + hi[1]), 0.5 * (lo[2] + hi[2])};
  double r2 = 0;
  for (int k = 1; k <= N; ++k) {
    const double dx = q[3 * k] - c[0], dy = q[3 * k + 1] - c[1], dz = q[3 * k + 2] - c[2];
    r2 = fmax(r2, dx * dx + dy * dy + dz * dz);
  }
  const double r = sqrt(r2) * (1.0 + 1e-9) + 1e-9;
  double* b = bounds + 4 * (size_t)j;
  b[0] = c[0], b[1] = c[1], b[2] = c[2], b[3] = (all_valid[j] != 0 && r == r) ? r : -1.0;  // NaN plans: never near
}

// K1 alone: the separating plane of every (own point, neighbour point) pair, out[i] = (n_f, b); NaN rows for
// coincident points.  The solver kernel inlines the same function; this entry exists so that the plane
// coefficients themselves can be compared with the reference's (agent_class.cpp:1152-1205).
__global__ void hdsm_planes_kernel(const hdsm_params prm, int n, const double* __restrict__ pc, const double* __restrict__ po,
                                   double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double a[3] = {pc[3 * i], pc[3 * i + 1], pc[3 * i + 2]}, b[3] = {po[3 * i], po[3 * i + 1], po[3 * i + 2]};
  double nf[3], off;
  if (!interagent_plane(prm, a, b, nf, off)) nf[0] = nf[1] = nf[2] = off = NAN;
  out[4 * i] = nf[0], out[4 * i + 1] = nf[1], out[4 * i + 2] = nf[2], out[4 * i + 3] = off;
}

// Read-back and failure fallback of one replanning step on the device, for callers that keep the swarm state in
// HBM (agent_class.cpp:962-987 read-back, :997-1019 fallback, :233-238 state advance with step_plan = 1):
//   usable result (OPTIMAL, or NODE_LIMIT with an incumbent)  traj_curr_ := traj, control_curr_ := ctrl, plan valid
//   otherwise, with a previous plan                            both shifted by one step, last element duplicated
//   otherwise                                                  nothing changes (no plan is published, :180-190)
// then state_curr_ := traj_curr_[1] for agents with a plan, and prev_pos := positions of traj_curr_ (the next
// call's prev_self_pos; state_ini_ repeated while there is no plan, :1103-1110).  One thread per (agent, entry).
__global__ void __launch_bounds__(256) hdsm_advance_kernel(int n, int N, const double* __restrict__ traj,
                                                             const double* __restrict__ ctrl, const hdsm_result* __restrict__ res,
                                                             double* __restrict__ traj_curr, double* __restrict__ ctrl_curr,
                                                             uint8_t* __restrict__ have_plan, double* __restrict__ x0,
                                                             double* __restrict__ prev_pos) {
  const int agent = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (agent >= n) return;
  const hdsm_result r = res[agent];
  const bool ok = r.status == HDSM_OPTIMAL || (r.status == HDSM_NODE_LIMIT && isfinite(r.obj));
  const bool had = have_plan[agent] != 0;
  double* tc = traj_curr + (size_t)agent * (N + 1) * 9;
  double* cc = ctrl_curr + (size_t)agent * N * 3;
  const int nt = (N + 1) * 9, nc = N * 3;
  if (ok) {
    for (int i = lane; i < nt; i += 32) tc[i] = traj[(size_t)agent * nt + i];
    for (int i = lane; i < nc; i += 32) cc[i] = ctrl[(size_t)agent * nc + i];
  } else if (had) {  // every lane reads its elements before any lane writes (the shift moves data across lanes)
    double tv[4], cv[1];
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const int i = lane + 32 * m;
      tv[m] = i < nt ? tc[i + 9 < nt ? i + 9 : i] : 0.0;
    }
    cv[0] = lane < nc ? cc[lane + 3 < nc ? lane + 3 : lane] : 0.0;
    double cv2 = lane + 32 < nc ? cc[lane + 35 < nc ? lane + 35 : lane + 32] : 0.0;
    __syncwarp();
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const int i = lane + 32 * m;
      if (i < nt) tc[i] = tv[m];
    }
    if (lane < nc) cc[lane] = cv[0];
    if (lane + 32 < nc) cc[lane + 32] = cv2;
  }
  __syncwarp();
  const bool have = ok || had;
  if (lane == 0) have_plan[agent] = have ? 1 : 0;
  if (have && lane < 9) x0[(size_t)agent * 9 + lane] = tc[9 + lane];
  __syncwarp();
  if (prev_pos)
    for (int i = lane; i < 3 * (N + 1); i += 32) {
      const int k = i / 3, a = i - 3 * k;
      prev_pos[(size_t)agent * 3 * (N + 1) + i] = have ? tc[k * 9 + a] : x0[(size_t)agent * 9 + a];
    }
}

}  // namespace hdsm
