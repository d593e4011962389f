/* CPU port of the HDSM per-agent trajectory optimisation (TEST / BASELINE INFRASTRUCTURE ONLY).
 *
 * Plain-C float64 restatement, used (a) by tests/ as a fast checker next to the NumPy oracle
 * (oracle/hdsm_oracle.py, which poses the problem un-condensed and is pinned against HiGHS), and
 * (b) by bench.py as the CPU baseline / "--impl reference" arm (kind "port": the reference's own
 * solver is Gurobi 10, closed source and absent - SURVEY.md 8(c)).  Nothing under
 * multi_agent_pkgs_b200/ links, loads or calls this file.
 *
 * Reference behaviour restated (paths relative to the reference checkout):
 *   bounds                multi_agent_planner/src/agent_class.cpp:2169-2188
 *   dynamics, terminal    agent_class.cpp:2071-2153, ModelODE :2155-2167
 *   objective, x0, rows   agent_class.cpp:858-941, :1071-1084
 *   inter-agent planes    agent_class.cpp:1086-1215, AddHyperplane :1217-1234
 *   binaries / one-hot    agent_class.cpp:2105-2113, :928-940
 *
 * Algorithm (same maths as the CUDA kernels, independent code): inputs condensed onto the null
 * space of the terminal equalities (per axis N-2 free variables), Mehrotra predictor-corrector on
 * the inequality-only QP with the barrier Hessian accumulated as one 3x3 block per horizon step,
 * dense Cholesky, and an exact depth-first branch and bound over per-step polytope candidate sets.
 * PARITY: pinned only through the NumPy oracle (KKT certificates + HiGHS); Gurobi unverified.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

#define MAXN 12
#define MAXNZ (MAXN - 2)
#define MAXNW (3 * MAXNZ)
#define MAXP 8
#define MAXQ (3 * MAXN - 2)
#define MAXW 8 /* widest search round */
#define FEAS_TOL 1e-6
#define PRUNE_MARGIN 1e-6
#define PRUNE_REL 1e-7 /* bound pruning, relative (the kernels' kPruneRel) */
#define STEP_FRAC 0.97
#define LOOSE_TOL 1e-6 /* accepted when the iteration limit is reached: still inside the 1e-6 KKT target */

enum { ORC_OPTIMAL = 0, ORC_INFEASIBLE = 1, ORC_MAX_ITER = 2, ORC_NUMERICAL = 3, ORC_NODE_LIMIT = 4,
       ORC_CUTOFF = 6 /* internal: the dual bound of a relaxation reached the incumbent, solve abandoned */ };

typedef struct {
  int32_t n_hor, poly_hor, rk4, max_iter, max_nodes, prune, width, warm_start;
  double dt, drag[3], r_u, r_x[6], r_n[6];
  double max_vel, min_acc_xy, max_acc_xy, min_acc_z, max_acc_z, max_jerk;
  double drone_radius, drone_z_offset, tilt, tol;
} orc_params;

typedef struct {
  int32_t status, iters, nodes, rows;
  double obj, kkt;
} orc_result;

typedef struct {
  int N, nz, nw, nq;
  double Apow[3][MAXN + 1][3][3];             /* Aa^k per axis */
  double cP[3][MAXN + 1][3], QP[3][MAXN + 1][MAXNZ];
  double cV[3][MAXN + 1][3], QV[3][MAXN + 1][MAXNZ];
  double cA[3][MAXN + 1][3], QA[3][MAXN + 1][MAXNZ];
  double Up[3][MAXN][3], Z[3][MAXN][MAXNZ];
  double cQ[3][MAXQ][3], EQ[3][MAXQ][MAXNZ], qlo[3][MAXQ], qhi[3][MAXQ];
  int qconst[3][MAXQ];
  int kp_const[MAXN + 1];
  double Hw[3][MAXNZ][MAXNZ], HwInv[3][MAXNZ][MAXNZ];
  double Ba[3][3];
} tables_t;

/* ------------------------------------------------------------------ small dense helpers */
/* In-place lower Cholesky.  A pivot that lost all its digits to cancellation (<= 1e-13 of the
 * original diagonal) marks a direction the barrier has pinned: that variable is frozen for this
 * solve (column zeroed, diagonal stored as 0) instead of aborting.  strict != 0: fail instead. */
static int chol(int n, double *A, int lda, int strict) {
  for (int j = 0; j < n; j++) {
    double orig = A[j * lda + j], d = orig;
    for (int k = 0; k < j; k++) d -= A[j * lda + k] * A[j * lda + k];
    if (!(d > 1e-13 * orig) || !(orig > 0.0)) {
      if (strict || !(orig == orig)) return 1;
      A[j * lda + j] = 0.0;
      for (int i = j + 1; i < n; i++) A[i * lda + j] = 0.0;
      continue;
    }
    d = sqrt(d);
    A[j * lda + j] = d;
    for (int i = j + 1; i < n; i++) {
      double v = A[i * lda + j];
      for (int k = 0; k < j; k++) v -= A[i * lda + k] * A[j * lda + k];
      A[i * lda + j] = v / d;
    }
  }
  return 0;
}
static void chol_solve(int n, const double *L, int lda, double *x) {
  for (int i = 0; i < n; i++) {
    double v = x[i];
    for (int k = 0; k < i; k++) v -= L[i * lda + k] * x[k];
    x[i] = L[i * lda + i] > 0 ? v / L[i * lda + i] : 0.0;
  }
  for (int i = n - 1; i >= 0; i--) {
    double v = x[i];
    for (int k = i + 1; k < n; k++) v -= L[k * lda + i] * x[k];
    x[i] = L[i * lda + i] > 0 ? v / L[i * lda + i] : 0.0;
  }
}

/* Experiment knob (ORC_LINSOLVE=inv): the linear algebra the CUDA kernel uses since round 2 - right-looking
 * LDL' that accumulates the inverse of the unit factor in the free upper triangle, the two triangular solves
 * replaced by two products - run inside this port to measure, on the CPU, what the explicit inverse does to
 * iteration counts and statuses before any GPU time is spent.  Default: the Cholesky above. */
static __thread double g_minratio;
static int ldlinv_factor(int n, double *A, int lda, double *invd) {
  double d0[MAXNW];
  for (int j = 0; j < n; j++) d0[j] = A[j * lda + j];
  for (int j = 0; j < n; j++) {
    const double djj = A[j * lda + j];
    if (!(d0[j] > 0.0)) return 1;
    const double inv = djj > 1e-13 * d0[j] ? 1.0 / djj : 0.0;
    invd[j] = inv;
    if (j == 0) g_minratio = 1.0;
    if (inv > 0 && djj / d0[j] < g_minratio) g_minratio = djj / d0[j];
    for (int i = j + 1; i < n; i++) {
      const double lij = A[i * lda + j] * inv;
      for (int k = 0; k <= i; k++) {
        if (k == j) A[k * lda + i] = -lij;                           /* Linv[i][j] */
        else if (k > j) A[i * lda + k] -= lij * A[k * lda + j];      /* trailing matrix */
        else A[k * lda + i] -= lij * A[k * lda + j];                 /* Linv[i][k] -= L_ij Linv[j][k] */
      }
    }
  }
  return 0;
}
static void ldlinv_solve(int n, const double *A, int lda, const double *invd, double *x) {
  double z[MAXNW];
  for (int i = 0; i < n; i++) {
    double v = x[i];
    for (int c = 0; c < i; c++) v += A[c * lda + i] * x[c];
    z[i] = v * invd[i];
  }
  for (int c = 0; c < n; c++) {
    double v = z[c];
    for (int i = c + 1; i < n; i++) v += A[c * lda + i] * z[i];
    x[c] = v;
  }
}
/* substitution on the same factor (lower triangle: unscaled columns c_ij = L_ij d_j) */
static void ldlinv_solve_subst(int n, const double *A, int lda, const double *invd, double *x) {
  double y[MAXNW];
  for (int i = 0; i < n; i++) {
    double v = x[i];
    for (int j = 0; j < i; j++) v -= A[i * lda + j] * y[j];
    y[i] = v * invd[i]; /* y = D^-1 z */
  }
  for (int i = n - 1; i >= 0; i--) {
    double v = y[i];
    for (int k = i + 1; k < n; k++) v -= A[k * lda + i] * invd[i] * x[k];
    x[i] = v;
  }
}
static int linsolve_inv(void) {
  const char *e = getenv("ORC_LINSOLVE");
  return e && !strcmp(e, "inv");
}

/* ModelODE restricted to one axis (agent_class.cpp:2155-2167): d/dt (p,v,a) = (v, a - c v, u) */
static void ode_axis(const double s[3], double u, double c, double out[3]) {
  out[0] = s[1];
  out[1] = s[2] - c * s[1];
  out[2] = u;
}
static void step_axis(const double s[3], double u, double c, double dt, int rk4, double out[3]) {
  double k1[3], k2[3], k3[3], k4[3], t[3];
  ode_axis(s, u, c, k1);
  if (rk4) { /* agent_class.cpp:2123-2139 */
    for (int j = 0; j < 3; j++) t[j] = s[j] + dt / 2 * k1[j];
    ode_axis(t, u, c, k2);
    for (int j = 0; j < 3; j++) t[j] = s[j] + dt / 2 * k2[j];
    ode_axis(t, u, c, k3);
    for (int j = 0; j < 3; j++) t[j] = s[j] + dt * k3[j];
    ode_axis(t, u, c, k4);
    for (int j = 0; j < 3; j++) out[j] = s[j] + dt * (k1[j] + 2 * k2[j] + 2 * k3[j] + k4[j]) / 6;
  } else {
    for (int j = 0; j < 3; j++) out[j] = s[j] + dt * k1[j];
  }
}

static int build_tables(const orc_params *P, tables_t *T) {
  int N = P->n_hor;
  if (N < 3 || N > MAXN || P->poly_hor < 1 || P->poly_hor > MAXP) return 1;
  memset(T, 0, sizeof(*T));
  T->N = N;
  T->nz = N - 2;
  T->nw = 3 * (N - 2);
  T->nq = 3 * N - 2;
  int nz = T->nz;
  for (int a = 0; a < 3; a++) {
    double Aa[3][3], Ba[3], e[3], o[3];
    for (int j = 0; j < 3; j++) {
      e[0] = e[1] = e[2] = 0;
      e[j] = 1;
      step_axis(e, 0.0, P->drag[a], P->dt, P->rk4, o);
      for (int i = 0; i < 3; i++) Aa[i][j] = o[i];
    }
    e[0] = e[1] = e[2] = 0;
    step_axis(e, 1.0, P->drag[a], P->dt, P->rk4, Ba);
    for (int i = 0; i < 3; i++) T->Ba[a][i] = Ba[i];
    /* G[k] (3 x N): s_k = Aa^k s0 + G[k] u */
    double G[MAXN + 1][3][MAXN];
    memset(G, 0, sizeof(G));
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) T->Apow[a][0][i][j] = (i == j);
    for (int k = 0; k < N; k++) {
      for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
          double v = 0;
          for (int l = 0; l < 3; l++) v += Aa[i][l] * T->Apow[a][k][l][j];
          T->Apow[a][k + 1][i][j] = v;
        }
      for (int i = 0; i < 3; i++)
        for (int c = 0; c < N; c++) {
          double v = 0;
          for (int l = 0; l < 3; l++) v += Aa[i][l] * G[k][l][c];
          G[k + 1][i][c] = v;
        }
      for (int i = 0; i < 3; i++) G[k + 1][i][k] += Ba[i];
    }
    /* terminal equalities E u = -(Aa^N s0)[1:3],  E = G[N][1:3] (2 x N) */
    double E[2][MAXN], EEt[2][2] = {{0, 0}, {0, 0}};
    for (int r = 0; r < 2; r++)
      for (int c = 0; c < N; c++) E[r][c] = G[N][1 + r][c];
    for (int r = 0; r < 2; r++)
      for (int q = 0; q < 2; q++)
        for (int c = 0; c < N; c++) EEt[r][q] += E[r][c] * E[q][c];
    double det = EEt[0][0] * EEt[1][1] - EEt[0][1] * EEt[1][0];
    if (fabs(det) < 1e-300) return 2;
    double inv[2][2] = {{EEt[1][1] / det, -EEt[0][1] / det}, {-EEt[1][0] / det, EEt[0][0] / det}};
    double Epinv[MAXN][2]; /* E^T (E E^T)^-1 */
    for (int c = 0; c < N; c++)
      for (int r = 0; r < 2; r++) Epinv[c][r] = E[0][c] * inv[0][r] + E[1][c] * inv[1][r];
    for (int c = 0; c < N; c++)
      for (int j = 0; j < 3; j++)
        T->Up[a][c][j] = -(Epinv[c][0] * T->Apow[a][N][1][j] + Epinv[c][1] * T->Apow[a][N][2][j]);
    /* orthonormal null-space basis: project unit vectors, modified Gram-Schmidt with pivoting */
    double Pn[MAXN][MAXN];
    for (int i = 0; i < N; i++)
      for (int j = 0; j < N; j++) Pn[i][j] = (i == j) - (Epinv[i][0] * E[0][j] + Epinv[i][1] * E[1][j]);
    int used[MAXN] = {0};
    for (int z = 0; z < nz; z++) {
      int best = -1;
      double bn = -1;
      double cand[MAXN][MAXN];
      for (int c = 0; c < N; c++) {
        if (used[c]) continue;
        for (int i = 0; i < N; i++) cand[c][i] = Pn[i][c];
        for (int rep = 0; rep < 2; rep++)
          for (int y = 0; y < z; y++) {
            double d = 0;
            for (int i = 0; i < N; i++) d += cand[c][i] * T->Z[a][i][y];
            for (int i = 0; i < N; i++) cand[c][i] -= d * T->Z[a][i][y];
          }
        double nn = 0;
        for (int i = 0; i < N; i++) nn += cand[c][i] * cand[c][i];
        if (nn > bn) bn = nn, best = c;
      }
      if (best < 0 || bn < 1e-20) return 3;
      used[best] = 1;
      bn = sqrt(bn);
      for (int i = 0; i < N; i++) T->Z[a][i][z] = cand[best][i] / bn;
    }
    /* affine maps of positions / velocities / accelerations in (s0, w) */
    for (int k = 0; k <= N; k++)
      for (int comp = 0; comp < 3; comp++) {
        double *cc = comp == 0 ? T->cP[a][k] : comp == 1 ? T->cV[a][k] : T->cA[a][k];
        double *qq = comp == 0 ? T->QP[a][k] : comp == 1 ? T->QV[a][k] : T->QA[a][k];
        for (int j = 0; j < 3; j++) {
          double v = T->Apow[a][k][comp][j];
          for (int c = 0; c < N; c++) v += G[k][comp][c] * T->Up[a][c][j];
          cc[j] = v;
        }
        for (int z = 0; z < nz; z++) {
          double v = 0, big = 0;
          for (int c = 0; c < N; c++) {
            v += G[k][comp][c] * T->Z[a][c][z];
            big = fmax(big, fabs(G[k][comp][c]));
          }
          qq[z] = (big == 0.0) ? 0.0 : v; /* structurally constant rows stay exactly zero */
        }
      }
    /* box quantities: jerk k=0..N-1, vel k=1..N-1, acc k=1..N-1 (agent_class.cpp:2083-2097) */
    double alo = a < 2 ? P->min_acc_xy : P->min_acc_z, ahi = a < 2 ? P->max_acc_xy : P->max_acc_z;
    int q = 0;
    for (int k = 0; k < N; k++, q++) {
      memcpy(T->cQ[a][q], T->Up[a][k], 3 * sizeof(double));
      memcpy(T->EQ[a][q], T->Z[a][k], nz * sizeof(double));
      T->qlo[a][q] = -P->max_jerk;
      T->qhi[a][q] = P->max_jerk;
    }
    for (int k = 1; k < N; k++, q++) {
      memcpy(T->cQ[a][q], T->cV[a][k], 3 * sizeof(double));
      memcpy(T->EQ[a][q], T->QV[a][k], nz * sizeof(double));
      T->qlo[a][q] = -P->max_vel;
      T->qhi[a][q] = P->max_vel;
    }
    for (int k = 1; k < N; k++, q++) {
      memcpy(T->cQ[a][q], T->cA[a][k], 3 * sizeof(double));
      memcpy(T->EQ[a][q], T->QA[a][k], nz * sizeof(double));
      T->qlo[a][q] = alo;
      T->qhi[a][q] = ahi;
    }
    for (q = 0; q < T->nq; q++) {
      double big = 0;
      for (int z = 0; z < nz; z++) big = fmax(big, fabs(T->EQ[a][q][z]));
      T->qconst[a][q] = big == 0.0;
    }
    /* constant Hessian block: 2 r_u Z'Z + 2 sum_i (w_p QP_i QP_i' + w_v QV_i QV_i')  (:870-883, :2098) */
    for (int y = 0; y < nz; y++)
      for (int z = 0; z < nz; z++) {
        double v = 0;
        for (int c = 0; c < N; c++) v += T->Z[a][c][y] * T->Z[a][c][z];
        v *= 2 * P->r_u;
        for (int i = 1; i <= N; i++) {
          const double *wt = i == N ? P->r_n : P->r_x;
          v += 2 * wt[a] * T->QP[a][i][y] * T->QP[a][i][z] + 2 * wt[3 + a] * T->QV[a][i][y] * T->QV[a][i][z];
        }
        T->Hw[a][y][z] = v;
      }
    double L[MAXNZ][MAXNZ];
    memcpy(L, T->Hw[a], sizeof(L));
    if (chol(nz, &L[0][0], MAXNZ, 1)) return 4;
    for (int c = 0; c < nz; c++) {
      double e2[MAXNZ] = {0};
      e2[c] = 1;
      chol_solve(nz, &L[0][0], MAXNZ, e2);
      for (int r = 0; r < nz; r++) T->HwInv[a][r][c] = e2[r];
    }
  }
  for (int k = 0; k <= N; k++) {
    double big = 0;
    for (int a = 0; a < 3; a++)
      for (int z = 0; z < nz; z++) big = fmax(big, fabs(T->QP[a][k][z]));
    T->kp_const[k] = big == 0.0;
  }
  return 0;
}

/* ------------------------------------------------------------------ inter-agent plane (agent_class.cpp:1152-1205) */
static void interagent_plane(const orc_params *P, const double pc[3], const double po[3], double nf[3], double *b) {
  double n[3] = {po[0] - pc[0], po[1] - pc[1], po[2] - pc[2]};
  double nrm = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
  double nn[3] = {n[0] / nrm, n[1] / nrm, n[2] / nrm};
  double mid[3] = {(pc[0] + po[0]) / 2, (pc[1] + po[1]) / 2, (pc[2] + po[2]) / 2};
  double cz = nn[2] > 1 ? 1 : (nn[2] < -1 ? -1 : nn[2]);
  double ang = M_PI_2 - fabs(acos(cz));
  double t = atan(P->drone_radius / P->drone_z_offset * tan(ang));
  double sd = hypot(P->drone_radius * cos(t), P->drone_z_offset * sin(t));
  double h = fmin(2 * sd, nrm) / 2;
  double pt[3] = {mid[0] - h * nn[0], mid[1] - h * nn[1], mid[2] - h * nn[2]};
  /* right = nn x (0,0,1) + nn x (0,1,0); up_final = nn x (0,1,0) */
  double c1[3] = {nn[1], -nn[0], 0.0};
  double c2[3] = {-nn[2], 0.0, nn[0]};
  for (int i = 0; i < 3; i++) nf[i] = P->tilt * (c1[i] + c2[i]) + P->tilt * c2[i] + nn[i];
  *b = nf[0] * pt[0] + nf[1] * pt[1] + nf[2] * pt[2];
}

/* ------------------------------------------------------------------ per-agent problem */
typedef struct { double n[3], b; int kp; } prow_t;

typedef struct {
  const tables_t *T;
  const orc_params *P;
  double s0[3][3];            /* per-axis (p, v, a) of x0 */
  double g[MAXNW], c0;        /* objective 1/2 w'Hw w + g'w + c0 */
  double pbar[MAXN + 1][3];   /* positions at w = 0 */
  double qbar[3][MAXQ];       /* box quantities at w = 0 */
  double plo[MAXN + 1][3], phi[MAXN + 1][3]; /* reachable box of p_k (interval propagation) */
  prow_t *rows;               /* IPM position rows (variable kp only) */
  double *s, *lam;
  int m, cap;
  double cutoff;              /* a relaxation whose dual bound reaches this value cannot improve the incumbent */
} prob_t;

static void setup_problem(prob_t *pb, const double x0[9], const double *ref /*[N][6]*/) {
  const tables_t *T = pb->T;
  const orc_params *P = pb->P;
  int N = T->N, nz = T->nz;
  for (int a = 0; a < 3; a++)
    for (int j = 0; j < 3; j++) pb->s0[a][j] = x0[3 * j + a];
  pb->c0 = 0;
  for (int a = 0; a < 3; a++) {
    const double *s0 = pb->s0[a];
    double *g = pb->g + a * nz;
    for (int z = 0; z < nz; z++) g[z] = 0;
    for (int k = 0; k < N; k++) { /* r_u |Up s0 + Z w|^2 */
      double up = T->Up[a][k][0] * s0[0] + T->Up[a][k][1] * s0[1] + T->Up[a][k][2] * s0[2];
      pb->c0 += P->r_u * up * up;
      for (int z = 0; z < nz; z++) g[z] += 2 * P->r_u * T->Z[a][k][z] * up;
    }
    for (int i = 1; i <= N; i++) {
      const double *wt = i == N ? P->r_n : P->r_x;
      double ep = T->cP[a][i][0] * s0[0] + T->cP[a][i][1] * s0[1] + T->cP[a][i][2] * s0[2] - ref[(i - 1) * 6 + a];
      double ev = T->cV[a][i][0] * s0[0] + T->cV[a][i][1] * s0[1] + T->cV[a][i][2] * s0[2] - ref[(i - 1) * 6 + 3 + a];
      pb->c0 += wt[a] * ep * ep + wt[3 + a] * ev * ev;
      for (int z = 0; z < nz; z++) g[z] += 2 * wt[a] * T->QP[a][i][z] * ep + 2 * wt[3 + a] * T->QV[a][i][z] * ev;
    }
    for (int k = 0; k <= N; k++)
      pb->pbar[k][a] = T->cP[a][k][0] * s0[0] + T->cP[a][k][1] * s0[1] + T->cP[a][k][2] * s0[2];
    for (int q = 0; q < T->nq; q++)
      pb->qbar[a][q] = T->cQ[a][q][0] * s0[0] + T->cQ[a][q][1] * s0[1] + T->cQ[a][q][2] * s0[2];
    /* reachable interval of (p,v,a)_k under the jerk / acc / vel boxes */
    double lo[3] = {s0[0], s0[1], s0[2]}, hi[3] = {s0[0], s0[1], s0[2]};
    double alo = a < 2 ? P->min_acc_xy : P->min_acc_z, ahi = a < 2 ? P->max_acc_xy : P->max_acc_z;
    pb->plo[0][a] = pb->phi[0][a] = s0[0];
    for (int k = 0; k < N; k++) {
      double nl[3], nh[3];
      for (int i = 0; i < 3; i++) {
        double l = 0, h = 0;
        for (int j = 0; j < 3; j++) {
          double c = T->Apow[a][1][i][j];
          l += c >= 0 ? c * lo[j] : c * hi[j];
          h += c >= 0 ? c * hi[j] : c * lo[j];
        }
        double bc = T->Ba[a][i];
        l += bc >= 0 ? -bc * P->max_jerk : bc * P->max_jerk;
        h += bc >= 0 ? bc * P->max_jerk : -bc * P->max_jerk;
        nl[i] = l;
        nh[i] = h;
      }
      if (k + 1 < N) { /* boxes hold for k = 1..N-1 */
        nl[1] = fmax(nl[1], -P->max_vel), nh[1] = fmin(nh[1], P->max_vel);
        nl[2] = fmax(nl[2], alo), nh[2] = fmin(nh[2], ahi);
        if (nl[1] > nh[1]) nl[1] = nh[1] = 0.5 * (nl[1] + nh[1]); /* infeasible boxes: keep a valid point, IPM decides */
        if (nl[2] > nh[2]) nl[2] = nh[2] = 0.5 * (nl[2] + nh[2]);
      }
      memcpy(lo, nl, sizeof(lo));
      memcpy(hi, nh, sizeof(hi));
      pb->plo[k + 1][a] = lo[0];
      pb->phi[k + 1][a] = hi[0];
    }
  }
}

/* can the row n.p_kp <= b ever be active inside the reachable box? */
static int row_reachable(const prob_t *pb, const double n[3], double b, int kp) {
  double mx = 0;
  for (int a = 0; a < 3; a++) mx += n[a] >= 0 ? n[a] * pb->phi[kp][a] : n[a] * pb->plo[kp][a];
  return !(mx <= b - PRUNE_MARGIN);
}

static void push_row(prob_t *pb, const double n[3], double b, int kp) {
  if (pb->m == pb->cap) {
    pb->cap = pb->cap ? 2 * pb->cap : 512;
    pb->rows = (prow_t *)realloc(pb->rows, pb->cap * sizeof(prow_t));
    pb->s = (double *)realloc(pb->s, pb->cap * sizeof(double));
    pb->lam = (double *)realloc(pb->lam, pb->cap * sizeof(double));
  }
  prow_t *r = &pb->rows[pb->m++];
  r->n[0] = n[0], r->n[1] = n[1], r->n[2] = n[2], r->b = b, r->kp = kp;
}

/* ------------------------------------------------------------------ Mehrotra predictor-corrector in w-space */
typedef struct {
  int status, iters;
  double obj, w[MAXNW], kkt;
} qp_out;

static void positions(const prob_t *pb, const double *w, double p[MAXN + 1][3], int with_bar) {
  const tables_t *T = pb->T;
  for (int k = 0; k <= T->N; k++)
    for (int a = 0; a < 3; a++) {
      double v = with_bar ? pb->pbar[k][a] : 0.0;
      for (int z = 0; z < T->nz; z++) v += T->QP[a][k][z] * w[a * T->nz + z];
      p[k][a] = v;
    }
}
static void quantities(const prob_t *pb, const double *w, double qv[3][MAXQ], int with_bar) {
  const tables_t *T = pb->T;
  for (int a = 0; a < 3; a++)
    for (int q = 0; q < T->nq; q++) {
      double v = with_bar ? pb->qbar[a][q] : 0.0;
      for (int z = 0; z < T->nz; z++) v += T->EQ[a][q][z] * w[a * T->nz + z];
      qv[a][q] = v;
    }
}

static void solve_qp(prob_t *pb, qp_out *out) {
  const tables_t *T = pb->T;
  const orc_params *P = pb->P;
  const int N = T->N, nz = T->nz, nw = T->nw, nq = T->nq, m = pb->m;
  double w[MAXNW], p[MAXN + 1][3], qv[3][MAXQ];
  double bs[3][MAXQ][2], bl[3][MAXQ][2]; /* box slacks / multipliers: [..][0] upper, [..][1] lower */
  int nbox = 0;
  /* constant box rows (v_1 under Euler): feasibility check only */
  for (int a = 0; a < 3; a++)
    for (int q = 0; q < nq; q++)
      if (T->qconst[a][q]) {
        if (pb->qbar[a][q] - T->qhi[a][q] > FEAS_TOL || T->qlo[a][q] - pb->qbar[a][q] > FEAS_TOL) {
          out->status = ORC_INFEASIBLE, out->iters = 0, out->obj = INFINITY;
          return;
        }
      } else
        nbox += 2;
  const int mtot = m + nbox;
  /* start: unconstrained minimiser, slacks pushed positive, centred multipliers */
  for (int a = 0; a < 3; a++)
    for (int r = 0; r < nz; r++) {
      double v = 0;
      for (int c = 0; c < nz; c++) v -= T->HwInv[a][r][c] * pb->g[a * nz + c];
      w[a * nz + r] = v;
    }
  positions(pb, w, p, 1);
  quantities(pb, w, qv, 1);
  for (int i = 0; i < m; i++) {
    const prow_t *r = &pb->rows[i];
    double sl = r->b - (r->n[0] * p[r->kp][0] + r->n[1] * p[r->kp][1] + r->n[2] * p[r->kp][2]);
    pb->s[i] = fmax(sl, 1.0);
    pb->lam[i] = 1.0 / pb->s[i];
  }
  for (int a = 0; a < 3; a++)
    for (int q = 0; q < nq; q++) {
      if (T->qconst[a][q]) continue;
      double sc = 0.05 * (T->qhi[a][q] - T->qlo[a][q]);
      bs[a][q][0] = fmax(T->qhi[a][q] - qv[a][q], sc);
      bs[a][q][1] = fmax(qv[a][q] - T->qlo[a][q], sc);
      bl[a][q][0] = 1.0 / bs[a][q][0];
      bl[a][q][1] = 1.0 / bs[a][q][1];
    }
  double gmax = 0;
  for (int i = 0; i < nw; i++) gmax = fmax(gmax, fabs(pb->g[i]));
  const double tol = P->tol;
  int it;
  for (it = 0; it <= P->max_iter; it++) {
    /* pass 1: residuals, barrier blocks */
    double M[MAXN + 1][6], Tk[MAXN + 1][3], Fk[MAXN + 1][3];
    double DQ[3][MAXQ], TQ[3][MAXQ], FQ[3][MAXQ];
    memset(M, 0, sizeof(M)), memset(Tk, 0, sizeof(Tk)), memset(Fk, 0, sizeof(Fk));
    double mu = 0, rcmax = 0, lamsl = 0, lamsum = 0;
    for (int i = 0; i < m; i++) {
      const prow_t *r = &pb->rows[i];
      const double *pk = p[r->kp];
      double sl = r->b - (r->n[0] * pk[0] + r->n[1] * pk[1] + r->n[2] * pk[2]);
      double rc = pb->s[i] - sl, d = pb->lam[i] / pb->s[i], t = d * rc, l = pb->lam[i];
      mu += pb->s[i] * l, rcmax = fmax(rcmax, fabs(rc)), lamsl += l * sl, lamsum += l;
      double *Mk = M[r->kp];
      Mk[0] += d * r->n[0] * r->n[0], Mk[1] += d * r->n[0] * r->n[1], Mk[2] += d * r->n[0] * r->n[2];
      Mk[3] += d * r->n[1] * r->n[1], Mk[4] += d * r->n[1] * r->n[2], Mk[5] += d * r->n[2] * r->n[2];
      for (int a = 0; a < 3; a++) Tk[r->kp][a] += t * r->n[a], Fk[r->kp][a] += l * r->n[a];
    }
    for (int a = 0; a < 3; a++)
      for (int q = 0; q < nq; q++) {
        DQ[a][q] = TQ[a][q] = FQ[a][q] = 0;
        if (T->qconst[a][q]) continue;
        double slu = T->qhi[a][q] - qv[a][q], sll = qv[a][q] - T->qlo[a][q];
        double rcu = bs[a][q][0] - slu, rcl = bs[a][q][1] - sll;
        double du = bl[a][q][0] / bs[a][q][0], dl = bl[a][q][1] / bs[a][q][1];
        mu += bs[a][q][0] * bl[a][q][0] + bs[a][q][1] * bl[a][q][1];
        rcmax = fmax(rcmax, fmax(fabs(rcu), fabs(rcl)) / (T->qhi[a][q] - T->qlo[a][q]));
        lamsl += bl[a][q][0] * slu + bl[a][q][1] * sll, lamsum += bl[a][q][0] + bl[a][q][1];
        DQ[a][q] = du + dl;
        TQ[a][q] = du * rcu - dl * rcl; /* lower row has coefficient -E */
        FQ[a][q] = bl[a][q][0] - bl[a][q][1];
      }
    mu /= mtot > 0 ? mtot : 1;
    /* gradient of the smooth part and dual residual */
    double hg[MAXNW], fv[MAXNW], rdmax = 0, fmaxv = 0, wf = 0;
    for (int a = 0; a < 3; a++)
      for (int r = 0; r < nz; r++) {
        double v = pb->g[a * nz + r];
        for (int c = 0; c < nz; c++) v += T->Hw[a][r][c] * w[a * nz + c];
        hg[a * nz + r] = v;
        double f = 0;
        for (int k = 0; k <= N; k++) f += Fk[k][a] * T->QP[a][k][r];
        for (int q = 0; q < nq; q++) f += FQ[a][q] * T->EQ[a][q][r];
        rdmax = fmax(rdmax, fabs(v + f)), fmaxv = fmax(fmaxv, fabs(f)), wf += w[a * nz + r] * f;
        fv[a * nz + r] = pb->g[a * nz + r] + f;
      }
    double obj = pb->c0;
    for (int i = 0; i < nw; i++) obj += 0.5 * w[i] * (hg[i] + pb->g[i]);
    if (getenv("ORC_TRACE") && atoi(getenv("ORC_TRACE")) > 1)
      fprintf(stderr, "   it %d rd %.2e rc %.2e mu %.2e obj %.9f lamsum %.2e\n", it, rdmax / (1 + gmax), rcmax, mu, obj, lamsum);
#define ACCEPT(TOL) (rdmax <= (TOL) * (1 + gmax) && rcmax <= (TOL) && mu <= 0.1 * (TOL) * fmax(1.0, fabs(obj)))
    if (ACCEPT(tol) || (it == P->max_iter && ACCEPT(LOOSE_TOL))) {
      out->status = ORC_OPTIMAL, out->iters = it, out->obj = obj, out->kkt = fmax(rdmax / (1 + gmax), fmax(rcmax, mu / fmax(1.0, fabs(obj))));
      memcpy(out->w, w, sizeof(double) * nw);
      return;
    }
    /* Farkas certificate of primal infeasibility: lam >= 0, C'lam ~ 0, d'lam < 0 */
    if (mtot > 0 && lamsum > 0) {
      double dl = (lamsl + wf) / lamsum;
      if (fmaxv / lamsum < 1e-9 * fmax(1.0, -dl * 1e3) && dl < -1e-7 && it >= 3) {
        out->status = ORC_INFEASIBLE, out->iters = it, out->obj = INFINITY;
        return;
      }
    }
    /* Lagrangian dual bound (weak duality, any lam >= 0): min_w L(w, lam) = c0 - 1/2 v'Hw^-1 v - d'lam with
     * v = g + C'lam and d'lam = lam'(d - Cw) + w'C'lam.  Once it reaches the incumbent the node cannot win. */
    if (pb->cutoff < INFINITY) {
      double quad = 0;
      for (int a = 0; a < 3; a++)
        for (int r = 0; r < nz; r++) {
          double hv = 0;
          for (int c = 0; c < nz; c++) hv += T->HwInv[a][r][c] * fv[a * nz + c];
          quad += fv[a * nz + r] * hv;
        }
      if (pb->c0 - 0.5 * quad - (lamsl + wf) >= pb->cutoff) {
        out->status = ORC_CUTOFF, out->iters = it, out->obj = INFINITY;
        return;
      }
    }
    if (it == P->max_iter) break;
    /* assemble K = Hw + sum_k QP_k' M_k QP_k + sum_q DQ EQ EQ' */
    double K[MAXNW][MAXNW];
    memset(K, 0, sizeof(K));
    for (int a = 0; a < 3; a++)
      for (int r = 0; r < nz; r++)
        for (int c = 0; c < nz; c++) {
          double v = T->Hw[a][r][c];
          for (int q = 0; q < nq; q++) v += DQ[a][q] * T->EQ[a][q][r] * T->EQ[a][q][c];
          K[a * nz + r][a * nz + c] = v;
        }
    static const int mi[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};
    for (int k = 0; k <= N; k++) {
      if (T->kp_const[k]) continue;
      for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++) {
          double mk = M[k][mi[a][b]];
          if (mk == 0.0) continue;
          for (int r = 0; r < nz; r++) {
            double f = mk * T->QP[a][k][r];
            for (int c = 0; c < nz; c++) K[a * nz + r][b * nz + c] += f * T->QP[b][k][c];
          }
        }
    }
    double invd[MAXNW];
    const int use_inv = linsolve_inv();
    if (use_inv ? ldlinv_factor(nw, &K[0][0], MAXNW, invd) : chol(nw, &K[0][0], MAXNW, 0)) {
      out->status = ORC_NUMERICAL, out->iters = it, out->obj = INFINITY;
      return;
    }
    /* predictor */
    double dw[MAXNW], dp[MAXN + 1][3], dq[3][MAXQ];
    for (int a = 0; a < 3; a++)
      for (int r = 0; r < nz; r++) {
        double v = -hg[a * nz + r];
        for (int k = 0; k <= N; k++) v -= Tk[k][a] * T->QP[a][k][r];
        for (int q = 0; q < nq; q++) v -= TQ[a][q] * T->EQ[a][q][r];
        dw[a * nz + r] = v;
      }
    if (use_inv) {
      if (getenv("ORC_LINLOG")) {
        double xs[MAXNW], xi[MAXNW], dmax = 0, xmax = 0;
        memcpy(xs, dw, sizeof(double) * nw), memcpy(xi, dw, sizeof(double) * nw);
        ldlinv_solve_subst(nw, &K[0][0], MAXNW, invd, xs);
        ldlinv_solve(nw, &K[0][0], MAXNW, invd, xi);
        for (int i = 0; i < nw; i++) dmax = fmax(dmax, fabs(xs[i] - xi[i])), xmax = fmax(xmax, fabs(xs[i]));
        FILE *f = fopen(getenv("ORC_LINLOG"), "a");
        fprintf(f, "%d %.3e %.3e %.3e %.3e\n", it, mu, g_minratio, dmax / (xmax > 0 ? xmax : 1), rdmax / (1 + gmax));
        fclose(f);
      }
      const char *thr = getenv("ORC_LINTHR");
      if (thr && g_minratio < atof(thr)) ldlinv_solve_subst(nw, &K[0][0], MAXNW, invd, dw);
      else ldlinv_solve(nw, &K[0][0], MAXNW, invd, dw);
    } else chol_solve(nw, &K[0][0], MAXNW, dw);
    positions(pb, dw, dp, 0);
    quantities(pb, dw, dq, 0);
    double alpha = 1.0, s_sl = 0, s_x = 0, s_dd = 0;
#define STEP_ROW(S, L, SL, CDW, CORR)                                   \
  {                                                                     \
    double rc_ = (S) - (SL);                                            \
    double ds_ = -rc_ - (CDW);                                          \
    double dl_ = -((S) * (L) + (CORR) + (L) * ds_) / (S);               \
    if (ds_ < 0) alpha = fmin(alpha, -(S) / ds_);                       \
    if (dl_ < 0) alpha = fmin(alpha, -(L) / dl_);                       \
    s_sl += (S) * (L), s_x += (S) * dl_ + (L) * ds_, s_dd += ds_ * dl_; \
  }
    for (int i = 0; i < m; i++) {
      const prow_t *r = &pb->rows[i];
      double sl = r->b - (r->n[0] * p[r->kp][0] + r->n[1] * p[r->kp][1] + r->n[2] * p[r->kp][2]);
      double cdw = r->n[0] * dp[r->kp][0] + r->n[1] * dp[r->kp][1] + r->n[2] * dp[r->kp][2];
      STEP_ROW(pb->s[i], pb->lam[i], sl, cdw, 0.0)
    }
    for (int a = 0; a < 3; a++)
      for (int q = 0; q < nq; q++) {
        if (T->qconst[a][q]) continue;
        STEP_ROW(bs[a][q][0], bl[a][q][0], T->qhi[a][q] - qv[a][q], dq[a][q], 0.0)
        STEP_ROW(bs[a][q][1], bl[a][q][1], qv[a][q] - T->qlo[a][q], -dq[a][q], 0.0)
      }
    double mu_aff = (s_sl + alpha * s_x + alpha * alpha * s_dd) / (mtot > 0 ? mtot : 1);
    if (getenv("ORC_TRACE") && atoi(getenv("ORC_TRACE")) > 2) fprintf(stderr, "      alpha_aff %.3e mu_aff %.3e\n", alpha, mu_aff);
    double sig = mu > 0 ? pow(fmax(mu_aff, 0.0) / mu, 3) : 0.0;
    double smu = sig * mu;
    /* corrector: T accumulations with the second-order term */
    double Tc[MAXN + 1][3], TQc[3][MAXQ];
    memset(Tc, 0, sizeof(Tc));
    for (int i = 0; i < m; i++) {
      const prow_t *r = &pb->rows[i];
      double sl = r->b - (r->n[0] * p[r->kp][0] + r->n[1] * p[r->kp][1] + r->n[2] * p[r->kp][2]);
      double cdw = r->n[0] * dp[r->kp][0] + r->n[1] * dp[r->kp][1] + r->n[2] * dp[r->kp][2];
      double S = pb->s[i], L = pb->lam[i], rc = S - sl, dsa = -rc - cdw, dla = -(S * L + L * dsa) / S;
      double t = (L * rc - (dsa * dla - smu)) / S;
      for (int a = 0; a < 3; a++) Tc[r->kp][a] += t * r->n[a];
    }
    for (int a = 0; a < 3; a++)
      for (int q = 0; q < nq; q++) {
        TQc[a][q] = 0;
        if (T->qconst[a][q]) continue;
        for (int sd = 0; sd < 2; sd++) {
          double S = bs[a][q][sd], L = bl[a][q][sd];
          double sl = sd == 0 ? T->qhi[a][q] - qv[a][q] : qv[a][q] - T->qlo[a][q];
          double cdw = sd == 0 ? dq[a][q] : -dq[a][q];
          double rc = S - sl, dsa = -rc - cdw, dla = -(S * L + L * dsa) / S;
          double t = (L * rc - (dsa * dla - smu)) / S;
          TQc[a][q] += sd == 0 ? t : -t;
        }
      }
    double dwc[MAXNW], dpc[MAXN + 1][3], dqc[3][MAXQ];
    for (int a = 0; a < 3; a++)
      for (int r = 0; r < nz; r++) {
        double v = -hg[a * nz + r];
        for (int k = 0; k <= N; k++) v -= Tc[k][a] * T->QP[a][k][r];
        for (int q = 0; q < nq; q++) v -= TQc[a][q] * T->EQ[a][q][r];
        dwc[a * nz + r] = v;
      }
    if (use_inv) {
      const char *thr = getenv("ORC_LINTHR");
      if (thr && g_minratio < atof(thr)) ldlinv_solve_subst(nw, &K[0][0], MAXNW, invd, dwc);
      else ldlinv_solve(nw, &K[0][0], MAXNW, invd, dwc);
    } else chol_solve(nw, &K[0][0], MAXNW, dwc);
    positions(pb, dwc, dpc, 0);
    quantities(pb, dwc, dqc, 0);
    /* final step length, then update */
    alpha = 1.0;
    s_sl = s_x = s_dd = 0;
    for (int i = 0; i < m; i++) {
      const prow_t *r = &pb->rows[i];
      double sl = r->b - (r->n[0] * p[r->kp][0] + r->n[1] * p[r->kp][1] + r->n[2] * p[r->kp][2]);
      double cdw = r->n[0] * dp[r->kp][0] + r->n[1] * dp[r->kp][1] + r->n[2] * dp[r->kp][2];
      double S = pb->s[i], L = pb->lam[i], rc = S - sl, dsa = -rc - cdw, dla = -(S * L + L * dsa) / S;
      double cdwc = r->n[0] * dpc[r->kp][0] + r->n[1] * dpc[r->kp][1] + r->n[2] * dpc[r->kp][2];
      STEP_ROW(S, L, sl, cdwc, dsa * dla - smu)
    }
    for (int a = 0; a < 3; a++)
      for (int q = 0; q < nq; q++) {
        if (T->qconst[a][q]) continue;
        for (int sd = 0; sd < 2; sd++) {
          double S = bs[a][q][sd], L = bl[a][q][sd];
          double sl = sd == 0 ? T->qhi[a][q] - qv[a][q] : qv[a][q] - T->qlo[a][q];
          double cdw = sd == 0 ? dq[a][q] : -dq[a][q], cdwc = sd == 0 ? dqc[a][q] : -dqc[a][q];
          double rc = S - sl, dsa = -rc - cdw, dla = -(S * L + L * dsa) / S;
          STEP_ROW(S, L, sl, cdwc, dsa * dla - smu)
        }
      }
    /* 0.97 of the way to the boundary: 0.995 leaves the blocking pair so far off the central path
     * that predictor and centring steps alternate without reducing mu on ~0.4% of the QPs */
    double al = fmin(1.0, STEP_FRAC * alpha);
    if (getenv("ORC_TRACE") && atoi(getenv("ORC_TRACE")) > 2) fprintf(stderr, "      sig %.3e alpha %.3e\n", sig, al);
    for (int i = 0; i < m; i++) {
      const prow_t *r = &pb->rows[i];
      double sl = r->b - (r->n[0] * p[r->kp][0] + r->n[1] * p[r->kp][1] + r->n[2] * p[r->kp][2]);
      double cdw = r->n[0] * dp[r->kp][0] + r->n[1] * dp[r->kp][1] + r->n[2] * dp[r->kp][2];
      double S = pb->s[i], L = pb->lam[i], rc = S - sl, dsa = -rc - cdw, dla = -(S * L + L * dsa) / S;
      double cdwc = r->n[0] * dpc[r->kp][0] + r->n[1] * dpc[r->kp][1] + r->n[2] * dpc[r->kp][2];
      double ds = -rc - cdwc, dl = -(S * L + (dsa * dla - smu) + L * ds) / S;
      pb->s[i] = S + al * ds;
      pb->lam[i] = L + al * dl;
    }
    for (int a = 0; a < 3; a++)
      for (int q = 0; q < nq; q++) {
        if (T->qconst[a][q]) continue;
        for (int sd = 0; sd < 2; sd++) {
          double S = bs[a][q][sd], L = bl[a][q][sd];
          double sl = sd == 0 ? T->qhi[a][q] - qv[a][q] : qv[a][q] - T->qlo[a][q];
          double cdw = sd == 0 ? dq[a][q] : -dq[a][q], cdwc = sd == 0 ? dqc[a][q] : -dqc[a][q];
          double rc = S - sl, dsa = -rc - cdw, dla = -(S * L + L * dsa) / S;
          double ds = -rc - cdwc, dl = -(S * L + (dsa * dla - smu) + L * ds) / S;
          bs[a][q][sd] = S + al * ds;
          bl[a][q][sd] = L + al * dl;
        }
      }
    for (int i = 0; i < nw; i++) w[i] += al * dwc[i];
    positions(pb, w, p, 1);
    quantities(pb, w, qv, 1);
    int bad = 0;
    for (int i = 0; i < nw; i++) bad |= !isfinite(w[i]);
    if (bad) {
      out->status = ORC_NUMERICAL, out->iters = it, out->obj = INFINITY;
      return;
    }
  }
  out->status = ORC_MAX_ITER, out->iters = it, out->obj = INFINITY;
  memcpy(out->w, w, sizeof(double) * nw);
}

/* ------------------------------------------------------------------ candidate-set rows (union hull) */
typedef struct { int n; double A[32][3], b[32]; } rowset_t;

static void set_rows(const double *pA, const double *pb_, const int32_t *prow_n, int rmax, unsigned mask, rowset_t *out) {
  int first = -1, cnt = 0;
  for (int j = 0; j < MAXP; j++)
    if (mask >> j & 1) {
      if (first < 0) first = j;
      cnt++;
    }
  out->n = 0;
  const double *A0 = pA + (size_t)first * rmax * 3, *b0 = pb_ + (size_t)first * rmax;
  if (cnt == 1) {
    for (int i = 0; i < prow_n[first]; i++) {
      memcpy(out->A[out->n], A0 + 3 * i, 3 * sizeof(double));
      out->b[out->n++] = b0[i];
    }
    return;
  }
  for (int i = 0; i < prow_n[first]; i++) {
    const double *a = A0 + 3 * i;
    int dup = 0;
    for (int r = 0; r < out->n; r++) dup |= out->A[r][0] == a[0] && out->A[r][1] == a[1] && out->A[r][2] == a[2];
    if (dup) continue;
    double bmax = -INFINITY;
    int ok = 1;
    for (int j = 0; j < MAXP && ok; j++) {
      if (!(mask >> j & 1)) continue;
      const double *Aj = pA + (size_t)j * rmax * 3, *bj = pb_ + (size_t)j * rmax;
      double bmin = INFINITY;
      for (int r = 0; r < prow_n[j]; r++)
        if (Aj[3 * r] == a[0] && Aj[3 * r + 1] == a[1] && Aj[3 * r + 2] == a[2]) bmin = fmin(bmin, bj[r]);
      if (bmin == INFINITY) ok = 0;
      bmax = fmax(bmax, bmin);
    }
    if (ok) {
      memcpy(out->A[out->n], a, 3 * sizeof(double));
      out->b[out->n++] = bmax;
    }
  }
}

static double seg_violation(const double *A, const double *b, int n, const double pa[3], const double pb_[3]) {
  double v = -INFINITY;
  for (int i = 0; i < n; i++) {
    v = fmax(v, A[3 * i] * pa[0] + A[3 * i + 1] * pa[1] + A[3 * i + 2] * pa[2] - b[i]);
    v = fmax(v, A[3 * i] * pb_[0] + A[3 * i + 1] * pb_[1] + A[3 * i + 2] * pb_[2] - b[i]);
  }
  return v;
}


/* ------------------------------------------------------------------ per-step dominance between cells
 * Cell A dominates cell B at step k when every point the segment (p_k, p_k+1) can reach inside B also lies in
 * A: then any trajectory that uses B at step k may use A instead at the same cost, and B is dropped from the
 * candidate set (the optimum value is unchanged; near-duplicate overlapping cells otherwise make the search
 * enumerate assignments that tie).  Sufficient test per row (n, b) of A and per variable point kp in {k, k+1}:
 * the row cannot be violated inside reach_box(kp) /\ bbox(B), or B holds the same normal at least as tight. */
static void cell_bbox(const double *A, const double *b, int n, double lo[3], double hi[3]) {
  for (int a = 0; a < 3; a++) lo[a] = -INFINITY, hi[a] = INFINITY;
  for (int r = 0; r < n; r++) {
    const double *v = A + 3 * r;
    int nzc = (v[0] != 0.0) + (v[1] != 0.0) + (v[2] != 0.0);
    if (nzc != 1) continue;
    int a = v[0] != 0.0 ? 0 : (v[1] != 0.0 ? 1 : 2);
    double x = b[r] / v[a];
    if (v[a] > 0) hi[a] = fmin(hi[a], x);
    else lo[a] = fmax(lo[a], x);
  }
}
static int cell_dominates(const prob_t *pb, const tables_t *T, int k, const double *AA, const double *bA, int nA,
                          const double *AB, const double *bB, int nB) {
  double lo[3], hi[3];
  cell_bbox(AB, bB, nB, lo, hi);
  for (int r = 0; r < nA; r++) {
    const double *n = AA + 3 * r;
    double same = INFINITY;
    for (int q = 0; q < nB; q++)
      if (AB[3 * q] == n[0] && AB[3 * q + 1] == n[1] && AB[3 * q + 2] == n[2]) same = fmin(same, bB[q]);
    if (same <= bA[r]) continue;
    for (int kp = k; kp <= k + 1; kp++) {
      if (T->kp_const[kp]) continue; /* a constant point was checked against both cells already */
      double mx = 0;
      for (int a = 0; a < 3; a++)
        mx += n[a] >= 0 ? n[a] * fmin(pb->phi[kp][a], hi[a]) : n[a] * fmax(pb->plo[kp][a], lo[a]);
      if (!(mx <= bA[r])) return 0;
    }
  }
  return 1;
}

/* ------------------------------------------------------------------ one agent */
static void solve_agent(const orc_params *P, const tables_t *T, int gid, int nb0, int nb1, const double *x0, const double *ref,
                        const double *pA, const double *pb_, const int32_t *prow_n, int rmax, const double *prev,
                        const double *all_pos, const uint8_t *all_valid, const int32_t *assign_in, double *traj,
                        double *ctrl, uint8_t *poly_used, int32_t *assign_out, orc_result *res) {
  const int N = T->N, nz = T->nz;
  prob_t pb;
  memset(&pb, 0, sizeof(pb));
  pb.T = T, pb.P = P;
  setup_problem(&pb, x0, ref);
  res->status = ORC_INFEASIBLE, res->iters = 0, res->nodes = 0, res->rows = 0, res->obj = INFINITY, res->kkt = INFINITY;
  for (int j = 0; j < P->poly_hor; j++) poly_used[j] = 0;
  for (int k = 0; k < N; k++) assign_out[k] = -1;
  int Peff = 0;
  while (Peff < P->poly_hor && prow_n[Peff] > 0) Peff++;
  /* inter-agent planes (always enforced: they are appended to every polytope, :1205) */
  int nplanes = 0;
  prow_t *planes = (prow_t *)malloc(sizeof(prow_t) * (size_t)N * (nb1 - nb0 > 0 ? nb1 - nb0 : 1) * 2);
  int infeasible = Peff == 0; /* sum over an empty set == 1 (:939-940) */
  for (int k = 0; k < N && !infeasible; k++)
    for (int j = nb0; j < nb1; j++) {
      if (j == gid || !all_valid[j]) continue;
      double nf[3], b;
      interagent_plane(P, prev + 3 * (k + 1), all_pos + ((size_t)j * (N + 1) + k + 1) * 3, nf, &b);
      for (int kp = k; kp <= k + 1; kp++) {
        if (T->kp_const[kp]) {
          if (nf[0] * pb.pbar[kp][0] + nf[1] * pb.pbar[kp][1] + nf[2] * pb.pbar[kp][2] - b > FEAS_TOL) infeasible = 1;
          continue;
        }
        if (P->prune && !row_reachable(&pb, nf, b, kp)) continue;
        prow_t *r = &planes[nplanes++];
        r->n[0] = nf[0], r->n[1] = nf[1], r->n[2] = nf[2], r->b = b, r->kp = kp;
      }
    }
  /* candidate sets; constant points filter them */
  unsigned root[MAXN];
  for (int k = 0; k < N && !infeasible; k++) {
    unsigned mask = 0;
    for (int j = 0; j < Peff; j++) {
      if (assign_in && assign_in[k] >= 0 && assign_in[k] != j) continue;
      int ok = 1;
      for (int kp = k; kp <= k + 1; kp++)
        if (T->kp_const[kp])
          for (int r = 0; r < prow_n[j]; r++) {
            const double *a = pA + ((size_t)j * rmax + r) * 3;
            if (a[0] * pb.pbar[kp][0] + a[1] * pb.pbar[kp][1] + a[2] * pb.pbar[kp][2] - pb_[(size_t)j * rmax + r] > FEAS_TOL) ok = 0;
          }
      if (ok) mask |= 1u << j;
    }
    if (!mask) infeasible = 1;
    root[k] = mask;
  }
  if (infeasible) {
    free(planes);
    return;
  }
  /* per-step dominance: drop a cell another candidate of the same step covers wherever the segment can be
   * (equal cells: the lower index stays) */
  for (int k = 0; k < N; k++)
    for (int B = Peff - 1; B >= 0; B--) {
      if (!(root[k] >> B & 1)) continue;
      for (int A = 0; A < Peff; A++) {
        if (A == B || !(root[k] >> A & 1)) continue;
        const double *AA = pA + (size_t)A * rmax * 3, *bA = pb_ + (size_t)A * rmax;
        const double *AB = pA + (size_t)B * rmax * 3, *bB = pb_ + (size_t)B * rmax;
        if (!cell_dominates(&pb, T, k, AA, bA, prow_n[A], AB, bB, prow_n[B])) continue;
        if (A > B && cell_dominates(&pb, T, k, AB, bB, prow_n[B], AA, bA, prow_n[A])) continue; /* equal cells */
        root[k] &= ~(1u << B);
        break;
      }
    }
  /* depth-first branch and bound over candidate sets */
  int cap = 4 * N * MAXP + 8, top = 0;
  unsigned(*stack)[MAXN] = malloc(sizeof(unsigned[MAXN]) * cap);
  double *sbound = malloc(sizeof(double) * cap);
  sbound[top] = -INFINITY;
  memcpy(stack[top++], root, sizeof(root));
  double best = INFINITY, bestw[MAXNW];
  int bestsig[MAXN], nodes = 0, iters = 0, exhausted = 1, maxrows = 0, anyfail = 0, lat_iters = 0;
  rowset_t rs;
  /* The search runs in rounds of up to `width` open nodes (width = 1: plain depth-first search).  The nodes of a
   * round are solved independently, all against the incumbent the round started with - on the GPU by the thread
   * blocks of one cluster at the same time - and then merged in a fixed order, so the result does not depend on
   * which of them finishes first: incumbents in pop order, then children pushed so that the first node's
   * children are on top of the stack. */
  const int width = P->width > 1 ? (P->width < MAXW ? P->width : MAXW) : 1;
  typedef struct {
    unsigned sets[MAXN];
    qp_out q;
    int full[MAXN], bk, no, order[MAXP];
    double p[MAXN + 1][3];
  } node_t;
  node_t *nd = (node_t *)malloc(sizeof(node_t) * width);
  while (top > 0 && exhausted) {
    /* ---- pop */
    int cnt = 0;
    while (cnt < width && top > 0) {
      if (nodes + cnt >= P->max_nodes) {
        if (cnt == 0) exhausted = 0;
        break;
      }
      --top;
      if (sbound[top] >= best - PRUNE_REL * fmax(1.0, fabs(best))) continue; /* the parent's optimum bounds this node: not solved, not counted */
      memcpy(nd[cnt++].sets, stack[top], sizeof(nd[0].sets));
    }
    if (cnt == 0) continue;
    /* ---- solve (independent; cutoff from the incumbent at the start of the round) */
    const double cutoff = best < INFINITY ? best - PRUNE_REL * fmax(1.0, fabs(best)) : INFINITY;
    int round_iters = 0;
    for (int c = 0; c < cnt; c++) {
      node_t *n = &nd[c];
      pb.m = 0;
      for (int i = 0; i < nplanes; i++) push_row(&pb, planes[i].n, planes[i].b, planes[i].kp);
      for (int k = 0; k < N; k++) {
        set_rows(pA, pb_, prow_n, rmax, n->sets[k], &rs);
        for (int kp = k; kp <= k + 1; kp++) {
          if (T->kp_const[kp]) continue;
          if (kp == k && k > 0 && n->sets[k - 1] == n->sets[k]) continue; /* same rows already on p_k from step k-1 */
          for (int r = 0; r < rs.n; r++)
            if (!P->prune || row_reachable(&pb, rs.A[r], rs.b[r], kp)) push_row(&pb, rs.A[r], rs.b[r], kp);
        }
      }
      if (pb.m > maxrows) maxrows = pb.m;
      pb.cutoff = cutoff;
      solve_qp(&pb, &n->q);
      nodes++;
      iters += n->q.iters;
      if (n->q.iters > round_iters) round_iters = n->q.iters;
      if (getenv("ORC_TRACE")) {
        fprintf(stderr, "node %d sets", nodes);
        for (int k = 0; k < N; k++) fprintf(stderr, " %x", n->sets[k]);
        fprintf(stderr, " rows %d status %d it %d obj %.6f best %.6f\n", pb.m, n->q.status, n->q.iters, n->q.obj, best);
      }
      n->bk = -1, n->no = 0;
      if (n->q.status != ORC_OPTIMAL) continue;
      /* coverage of every segment by one member of its candidate set; the uncovered step that is farthest from all
       * of its candidates is the one to branch on (~4x fewer nodes than "first uncovered" on the config-2 loop) */
      positions(&pb, n->q.w, n->p, 1);
      double viol[MAXP], bkv = 0;
      for (int k = 0; k < N; k++) {
        n->full[k] = -1;
        double v[MAXP];
        for (int j = 0; j < Peff; j++) {
          v[j] = INFINITY;
          if (n->sets[k] >> j & 1) {
            v[j] = seg_violation(pA + (size_t)j * rmax * 3, pb_ + (size_t)j * rmax, prow_n[j], n->p[k], n->p[k + 1]);
            if (n->full[k] < 0 && v[j] <= 1e-7) n->full[k] = j;
          }
        }
        double vmin = INFINITY;
        for (int j = 0; j < Peff; j++) vmin = fmin(vmin, v[j]);
        if (n->full[k] < 0 && (n->bk < 0 || vmin > bkv)) {
          bkv = vmin;
          n->bk = k;
          n->no = 0;
          for (int j = 0; j < Peff; j++)
            if (n->sets[k] >> j & 1) n->order[n->no] = j, viol[n->no++] = v[j];
          for (int x = 1; x < n->no; x++) /* insertion sort by (violation, index) */
            for (int y = x; y > 0 && viol[y] < viol[y - 1]; y--) {
              double tv = viol[y];
              viol[y] = viol[y - 1], viol[y - 1] = tv;
              int to = n->order[y];
              n->order[y] = n->order[y - 1], n->order[y - 1] = to;
            }
        }
      }
    }
    lat_iters += round_iters;
    const int first_round = nodes == cnt; /* this round was the root alone */
    /* ---- merge 1: incumbents, in pop order */
    for (int c = 0; c < cnt; c++) {
      node_t *n = &nd[c];
      if (n->q.status != ORC_OPTIMAL) {
        if (n->q.status != ORC_INFEASIBLE && n->q.status != ORC_CUTOFF) anyfail = n->q.status;
        continue;
      }
      if (n->bk >= 0 || n->q.obj >= best - PRUNE_REL * fmax(1.0, fabs(best))) continue;
      best = n->q.obj;
      memcpy(bestw, n->q.w, sizeof(bestw));
      memcpy(bestsig, n->full, sizeof(n->full));
      res->kkt = n->q.kkt;
    }
    /* ---- merge 2: children, last node first so that the first node's children end up on top */
    for (int c = cnt - 1; c >= 0 && exhausted; c--) {
      node_t *n = &nd[c];
      if (n->q.status != ORC_OPTIMAL || n->bk < 0 || n->no <= 1) continue;
      if (n->q.obj >= best - PRUNE_REL * fmax(1.0, fabs(best))) continue;
      /* Split the candidates of step bk, ordered by violation, in two halves (least violated explored first).
       * A half whose hull still contains the segment of the node optimum would be solved to the very same point
       * and then split again on the same step: that solve is skipped and the half is split right away. */
      struct { int x0, x1; } work[2 * MAXP];
      int nw = 0;
      const int bk = n->bk, h = (n->no + 1) / 2;
      work[nw].x0 = 0, work[nw++].x1 = h;   /* processed last-in first-out: the upper half is pushed first */
      work[nw].x0 = h, work[nw++].x1 = n->no;
      while (nw > 0) {
        const int x0 = work[--nw].x0, x1 = work[nw].x1;
        unsigned m = 0;
        for (int x = x0; x < x1; x++) m |= 1u << n->order[x];
        int contains = 0;
        if (x1 - x0 > 1) {
          set_rows(pA, pb_, prow_n, rmax, m, &rs);
          contains = seg_violation(&rs.A[0][0], rs.b, rs.n, n->p[bk], n->p[bk + 1]) <= 1e-7; /* the coverage tolerance */
        }
        if (contains) {
          const int hh = (x1 - x0 + 1) / 2;
          work[nw].x0 = x0, work[nw++].x1 = x0 + hh;
          work[nw].x0 = x0 + hh, work[nw++].x1 = x1;
          continue;
        }
        if (top + 1 > cap) {
          exhausted = 0; /* stack full: the optimum stays unproven */
          break;
        }
        memcpy(stack[top], n->sets, sizeof(n->sets));
        sbound[top] = n->q.obj;
        stack[top++][bk] = m;
      }
    }
    /* ---- warm start (optional): when the root has branched, the previous plan shifted by one step names a
     * cell per step - the cell of the root's candidate set the segment (prev[k+1], prev[k+2]) lies deepest in, if it lies in
     * one; that assignment goes on top of the stack, so the next round solves it first and an incumbent exists early */
    if (P->warm_start && first_round && top > 0 && exhausted && top + 1 <= cap) {
      unsigned hs[MAXN];
      int okh = 1;
      for (int k = 0; k < N && okh; k++) {
        const double *a0 = prev + 3 * (k + 1 <= N ? k + 1 : N), *a1 = prev + 3 * (k + 2 <= N ? k + 2 : N);
        int bj = -1;
        double bv = INFINITY;
        for (int j = 0; j < Peff; j++) {
          if (!(root[k] >> j & 1)) continue;
          double v = seg_violation(pA + (size_t)j * rmax * 3, pb_ + (size_t)j * rmax, prow_n[j], a0, a1);
          if (v < bv) bv = v, bj = j;
        }
        if (bj < 0 || bv > 1e-7) okh = 0;
        else hs[k] = 1u << bj;
      }
      if (okh) {
        memcpy(stack[top], hs, sizeof(hs));
        sbound[top++] = nd[0].q.obj;
      }
    }
  }
  free(nd);
  if (getenv("ORC_LAT")) maxrows = lat_iters; /* experiment: report the critical-path iterations in `rows` */
  res->nodes = nodes, res->iters = iters, res->rows = maxrows;
  if (best < INFINITY) {
    res->status = (exhausted && !anyfail) ? ORC_OPTIMAL : ORC_NODE_LIMIT; /* a lost node leaves the optimum unproven */
    res->obj = best;
    /* read-back (agent_class.cpp:962-987): u = Up s0 + Z w, states by the affine maps */
    for (int k = 0; k <= N; k++)
      for (int a = 0; a < 3; a++) {
        double pp = pb.pbar[k][a], vv = 0, aa = 0;
        for (int j = 0; j < 3; j++) vv += T->cV[a][k][j] * pb.s0[a][j], aa += T->cA[a][k][j] * pb.s0[a][j];
        for (int z = 0; z < nz; z++) {
          pp += T->QP[a][k][z] * bestw[a * nz + z];
          vv += T->QV[a][k][z] * bestw[a * nz + z];
          aa += T->QA[a][k][z] * bestw[a * nz + z];
        }
        traj[k * 9 + a] = pp, traj[k * 9 + 3 + a] = vv, traj[k * 9 + 6 + a] = aa;
      }
    for (int j = 0; j < 9; j++) traj[j] = x0[j];
    for (int j = 3; j < 9; j++) traj[N * 9 + j] = 0.0; /* fixed variables (:2078-2081) */
    for (int k = 0; k < N; k++)
      for (int a = 0; a < 3; a++) {
        double u = 0;
        for (int j = 0; j < 3; j++) u += T->Up[a][k][j] * pb.s0[a][j];
        for (int z = 0; z < nz; z++) u += T->Z[a][k][z] * bestw[a * nz + z];
        ctrl[k * 3 + a] = u;
      }
    for (int k = 0; k < N; k++) assign_out[k] = bestsig[k], poly_used[bestsig[k]] = 1;
  } else {
    res->status = !exhausted ? ORC_NODE_LIMIT : (anyfail ? anyfail : ORC_INFEASIBLE);
  }
  free(stack);
  free(sbound);
  free(planes);
  free(pb.rows), free(pb.s), free(pb.lam);
}

/* ------------------------------------------------------------------ batch entry (mirrors hdsm_solve_batch's arrays) */
typedef struct {
  const orc_params *P;
  const tables_t *T;
  int n_local, rmax, n_rob;
  const int32_t *global_id, *nbr_begin, *nbr_end, *poly_rows, *assign_in;
  const double *x0, *ref, *poly_A, *poly_b, *prev, *all_pos;
  const uint8_t *all_valid;
  double *traj, *ctrl;
  uint8_t *poly_used;
  int32_t *assign_out;
  orc_result *res;
  int next; /* work counter, claimed in chunks */
} job_t;

static void *worker(void *arg) {
  job_t *J = (job_t *)arg;
  const int N = J->P->n_hor, PH = J->P->poly_hor, rmax = J->rmax;
  for (;;) {
    int beg = __atomic_fetch_add(&J->next, 4, __ATOMIC_RELAXED);
    if (beg >= J->n_local) break;
    int end = beg + 4 < J->n_local ? beg + 4 : J->n_local;
    for (int i = beg; i < end; i++) {
      int nb0 = J->nbr_begin ? J->nbr_begin[i] : 0, nb1 = J->nbr_end ? J->nbr_end[i] : J->n_rob;
      solve_agent(J->P, J->T, J->global_id[i], nb0, nb1, J->x0 + (size_t)i * 9, J->ref + (size_t)i * N * 6,
                  J->poly_A + (size_t)i * PH * rmax * 3, J->poly_b + (size_t)i * PH * rmax,
                  J->poly_rows + (size_t)i * PH, rmax, J->prev + (size_t)i * (N + 1) * 3, J->all_pos, J->all_valid,
                  J->assign_in ? J->assign_in + (size_t)i * N : NULL, J->traj + (size_t)i * (N + 1) * 9,
                  J->ctrl + (size_t)i * N * 3, J->poly_used + (size_t)i * PH, J->assign_out + (size_t)i * N,
                  J->res + i);
    }
  }
  return NULL;
}

int orc_max_threads(void) {
  long n = sysconf(_SC_NPROCESSORS_ONLN);
  return n > 0 ? (int)n : 1;
}

int orc_solve_batch(const orc_params *P, int n_local, const int32_t *global_id, const int32_t *nbr_begin,
                    const int32_t *nbr_end, const double *x0, const double *ref, const double *poly_A,
                    const double *poly_b, const int32_t *poly_rows, int rmax, const double *prev_self_pos,
                    const double *all_pos, const uint8_t *all_valid, int n_rob, const int32_t *assign_in,
                    double *traj, double *ctrl, uint8_t *poly_used, int32_t *assign_out, orc_result *res,
                    int n_threads) {
  tables_t *T = (tables_t *)malloc(sizeof(tables_t));
  int rc = build_tables(P, T);
  if (rc) {
    free(T);
    return rc;
  }
  job_t J = {P, T, n_local, rmax, n_rob, global_id, nbr_begin, nbr_end, poly_rows, assign_in, x0, ref, poly_A, poly_b,
             prev_self_pos, all_pos, all_valid, traj, ctrl, poly_used, assign_out, res, 0};
  if (n_threads <= 0) n_threads = orc_max_threads();
  if (n_threads > (n_local + 3) / 4) n_threads = (n_local + 3) / 4;
  if (n_threads < 1) n_threads = 1;
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * n_threads);
  for (int t = 1; t < n_threads; t++) pthread_create(&th[t], NULL, worker, &J);
  worker(&J);
  for (int t = 1; t < n_threads; t++) pthread_join(th[t], NULL);
  free(th);
  free(T);
  return 0;
}

/* planes only: lets tests compare K1 against the NumPy restatement */
void orc_plane(const orc_params *P, const double *pc, const double *po, double *nf_b) {
  interagent_plane(P, pc, po, nf_b, nf_b + 3);
}
