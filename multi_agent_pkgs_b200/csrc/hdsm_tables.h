// Precomputed, parameter-only tables of the condensed trajectory QP (host builds, device reads).
//
// The reference poses the problem in all 9(N+1)+3N variables with the dynamics as equality rows
// (agent_class.cpp:2071-2153).  Here the states are eliminated with the (per-axis, 3x3) discrete
// dynamics and the six terminal equalities v_N = a_N = 0 (:2078-2081) are eliminated with an
// orthonormal null-space basis, leaving nz = N-2 free variables per axis:
//
//     u^a = Up^a s0^a + Z^a w^a          s0^a = (p, v, a) of axis a at k = 0
//     p_k^a = cP^a[k] . s0^a + QP^a[k] . w^a     (same for velocity cV/QV and acceleration cA/QA)
//
// Everything below depends on hdsm_params only, never on per-agent data.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#include "../../include/hdsm.h"

#if defined(__CUDACC__)
#define HDSM_HD __host__ __device__
#else
#define HDSM_HD
#endif

namespace hdsm {

constexpr int kMaxN = HDSM_MAX_HOR;
constexpr int kMaxNZ = kMaxN - 2;
constexpr int kMaxQ = 3 * kMaxN - 2;  // box quantities per axis: N jerks, N-1 velocities, N-1 accelerations
constexpr int kMaxP = HDSM_MAX_POLY;

struct Tables {
  int N, nz, nw, nq, nkp;
  int kp_of_slot[kMaxN + 1];  // variable position steps in ascending order
  int slot_of_kp[kMaxN + 1];  // -1: p_kp does not depend on w (rows on it are constants)
  int qconst[3][kMaxQ];       // box quantity independent of w (only checked, never in the QP)
  double cP[3][kMaxN + 1][3], cV[3][kMaxN + 1][3], cA[3][kMaxN + 1][3];
  double QP[3][kMaxN + 1][kMaxNZ], QV[3][kMaxN + 1][kMaxNZ], QA[3][kMaxN + 1][kMaxNZ];
  double Up[3][kMaxN][3], Z[3][kMaxN][kMaxNZ];
  double cQ[3][kMaxQ][3], EQ[3][kMaxQ][kMaxNZ], qlo[3][kMaxQ], qhi[3][kMaxQ];
  double Hw[3][kMaxNZ][kMaxNZ];     // constant Hessian block of axis a (objective 1/2 w'Hw w + g'w + c0)
  double HwInv[3][kMaxNZ][kMaxNZ];  // for the unconstrained-minimiser starting point
  double A1[3][3][3], B1[3][3];     // one-step per-axis dynamics (reachable-box propagation)
  hdsm_params prm;
};

namespace detail {

// d/dt (p, v, a) = (v, a - c v, u): ModelODE (agent_class.cpp:2155-2167) restricted to one axis.
inline void axis_rhs(const double s[3], double u, double c, double o[3]) {
  o[0] = s[1];
  o[1] = s[2] - c * s[1];
  o[2] = u;
}

// One integration step exactly as CreateGurobiModel writes it (:2117-2151).
inline void axis_step(const double s[3], double u, const hdsm_params& P, int axis, double o[3]) {
  const double c = P.drag[axis], dt = P.dt;
  double k1[3];
  axis_rhs(s, u, c, k1);
  if (!P.rk4) {
    for (int i = 0; i < 3; ++i) o[i] = s[i] + dt * k1[i];
    return;
  }
  double k2[3], k3[3], k4[3], t[3];
  for (int i = 0; i < 3; ++i) t[i] = s[i] + (dt / 2) * k1[i];
  axis_rhs(t, u, c, k2);
  for (int i = 0; i < 3; ++i) t[i] = s[i] + (dt / 2) * k2[i];
  axis_rhs(t, u, c, k3);
  for (int i = 0; i < 3; ++i) t[i] = s[i] + dt * k3[i];
  axis_rhs(t, u, c, k4);
  for (int i = 0; i < 3; ++i) o[i] = s[i] + dt * ((k1[i] + 2 * k2[i] + 2 * k3[i] + k4[i]) / 6);
}

// Orthonormal basis of null(E), E is 2 x n: Householder QR of E^T, last n-2 columns of Q.
inline bool null_space(const double E[2][kMaxN], int n, double Zout[kMaxN][kMaxNZ]) {
  double R[kMaxN][2], v[2][kMaxN], beta[2];
  for (int i = 0; i < n; ++i) R[i][0] = E[0][i], R[i][1] = E[1][i];
  for (int j = 0; j < 2; ++j) {
    double nrm = 0;
    for (int i = j; i < n; ++i) nrm += R[i][j] * R[i][j];
    nrm = std::sqrt(nrm);
    if (nrm < 1e-300) return false;
    const double alpha = R[j][j] >= 0 ? -nrm : nrm;
    for (int i = 0; i < n; ++i) v[j][i] = i < j ? 0.0 : R[i][j];
    v[j][j] -= alpha;
    double vv = 0;
    for (int i = j; i < n; ++i) vv += v[j][i] * v[j][i];
    beta[j] = vv > 0 ? 2.0 / vv : 0.0;
    for (int c = j; c < 2; ++c) {
      double d = 0;
      for (int i = j; i < n; ++i) d += v[j][i] * R[i][c];
      for (int i = j; i < n; ++i) R[i][c] -= beta[j] * d * v[j][i];
    }
  }
  if (std::fabs(R[1][1]) < 1e-14 * std::fabs(R[0][0])) return false;  // terminal rows dependent
  for (int z = 0; z < n - 2; ++z) {  // Q e_{z+2} = H0 H1 e_{z+2}
    double q[kMaxN];
    for (int i = 0; i < n; ++i) q[i] = i == z + 2 ? 1.0 : 0.0;
    for (int j = 1; j >= 0; --j) {
      double d = 0;
      for (int i = j; i < n; ++i) d += v[j][i] * q[i];
      for (int i = j; i < n; ++i) q[i] -= beta[j] * d * v[j][i];
    }
    for (int i = 0; i < n; ++i) Zout[i][z] = q[i];
  }
  return true;
}

inline bool cholesky_inverse(const double H[kMaxNZ][kMaxNZ], int n, double Hinv[kMaxNZ][kMaxNZ]) {
  double L[kMaxNZ][kMaxNZ] = {};
  for (int j = 0; j < n; ++j) {
    double d = H[j][j];
    for (int k = 0; k < j; ++k) d -= L[j][k] * L[j][k];
    if (!(d > 0)) return false;
    L[j][j] = std::sqrt(d);
    for (int i = j + 1; i < n; ++i) {
      double s = H[i][j];
      for (int k = 0; k < j; ++k) s -= L[i][k] * L[j][k];
      L[i][j] = s / L[j][j];
    }
  }
  for (int c = 0; c < n; ++c) {
    double x[kMaxNZ];
    for (int i = 0; i < n; ++i) {
      double s = i == c ? 1.0 : 0.0;
      for (int k = 0; k < i; ++k) s -= L[i][k] * x[k];
      x[i] = s / L[i][i];
    }
    for (int i = n - 1; i >= 0; --i) {
      double s = x[i];
      for (int k = i + 1; k < n; ++k) s -= L[k][i] * x[k];
      x[i] = s / L[i][i];
    }
    for (int i = 0; i < n; ++i) Hinv[i][c] = x[i];
  }
  return true;
}

}  // namespace detail

// Returns 0 on success.
inline int build_tables(const hdsm_params& P, Tables& T) {
  const int N = P.n_hor;
  // flat polytope-row indices are stored in a byte with 255 as the padding marker: poly_hor * Rmax <= 255
  if (N < HDSM_MIN_HOR || N > kMaxN || P.poly_hor < 1 || P.poly_hor > kMaxP || P.max_rows_per_poly < 1 ||
      P.max_rows_per_poly > 32 || P.poly_hor * P.max_rows_per_poly > 255 || !(P.dt > 0) || !(P.drone_z_offset > 0) ||
      !(P.max_jerk > 0))
    return 1;
  std::memset(&T, 0, sizeof(T));
  T.prm = P;
  T.N = N;
  T.nz = N - 2;
  T.nw = 3 * T.nz;
  T.nq = 3 * N - 2;
  const int nz = T.nz;
  for (int a = 0; a < 3; ++a) {
    // one-step maps from unit vectors (the step is linear and homogeneous)
    for (int j = 0; j < 3; ++j) {
      double e[3] = {0, 0, 0}, o[3];
      e[j] = 1;
      detail::axis_step(e, 0.0, P, a, o);
      for (int i = 0; i < 3; ++i) T.A1[a][i][j] = o[i];
    }
    {
      double e[3] = {0, 0, 0};
      detail::axis_step(e, 1.0, P, a, T.B1[a]);
    }
    // Phi[k] = A1^k,  G[k] (3 x N) input-to-state
    double Phi[kMaxN + 1][3][3] = {}, G[kMaxN + 1][3][kMaxN] = {};
    for (int i = 0; i < 3; ++i) Phi[0][i][i] = 1;
    for (int k = 0; k < N; ++k) {
      for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j)
          for (int l = 0; l < 3; ++l) Phi[k + 1][i][j] += T.A1[a][i][l] * Phi[k][l][j];
        for (int c = 0; c < k; ++c)
          for (int l = 0; l < 3; ++l) G[k + 1][i][c] += T.A1[a][i][l] * G[k][l][c];
        G[k + 1][i][k] = T.B1[a][i];
      }
    }
    // terminal rows  E u = -(Phi[N] s0)[1:3]
    double E[2][kMaxN], EEt[2][2] = {};
    for (int r = 0; r < 2; ++r)
      for (int c = 0; c < N; ++c) E[r][c] = G[N][1 + r][c];
    for (int r = 0; r < 2; ++r)
      for (int q = 0; q < 2; ++q)
        for (int c = 0; c < N; ++c) EEt[r][q] += E[r][c] * E[q][c];
    const double det = EEt[0][0] * EEt[1][1] - EEt[0][1] * EEt[1][0];
    if (!(std::fabs(det) > 0)) return 2;
    const double I2[2][2] = {{EEt[1][1] / det, -EEt[0][1] / det}, {-EEt[1][0] / det, EEt[0][0] / det}};
    for (int c = 0; c < N; ++c) {
      const double p0 = E[0][c] * I2[0][0] + E[1][c] * I2[1][0];  // row c of E^T (E E^T)^-1
      const double p1 = E[0][c] * I2[0][1] + E[1][c] * I2[1][1];
      for (int j = 0; j < 3; ++j) T.Up[a][c][j] = -(p0 * Phi[N][1][j] + p1 * Phi[N][2][j]);
    }
    if (!detail::null_space(E, N, T.Z[a])) return 3;
    // affine maps of p, v, a
    for (int k = 0; k <= N; ++k)
      for (int comp = 0; comp < 3; ++comp) {
        double* cc = comp == 0 ? T.cP[a][k] : comp == 1 ? T.cV[a][k] : T.cA[a][k];
        double* qq = comp == 0 ? T.QP[a][k] : comp == 1 ? T.QV[a][k] : T.QA[a][k];
        bool structural_zero = true;
        for (int c = 0; c < N; ++c) structural_zero &= G[k][comp][c] == 0.0;
        for (int j = 0; j < 3; ++j) {
          double s = Phi[k][comp][j];
          for (int c = 0; c < N; ++c) s += G[k][comp][c] * T.Up[a][c][j];
          cc[j] = s;
        }
        for (int z = 0; z < nz; ++z) {
          double s = 0;
          for (int c = 0; c < N; ++c) s += G[k][comp][c] * T.Z[a][c][z];
          qq[z] = structural_zero ? 0.0 : s;
        }
      }
    // box quantities (agent_class.cpp:2083-2097 with bounds :2179-2186)
    const double alo = a < 2 ? P.min_acc_xy : P.min_acc_z, ahi = a < 2 ? P.max_acc_xy : P.max_acc_z;
    int q = 0;
    for (int k = 0; k < N; ++k, ++q) {
      std::memcpy(T.cQ[a][q], T.Up[a][k], sizeof(double) * 3);
      std::memcpy(T.EQ[a][q], T.Z[a][k], sizeof(double) * nz);
      T.qlo[a][q] = -P.max_jerk, T.qhi[a][q] = P.max_jerk;
    }
    for (int k = 1; k < N; ++k, ++q) {
      std::memcpy(T.cQ[a][q], T.cV[a][k], sizeof(double) * 3);
      std::memcpy(T.EQ[a][q], T.QV[a][k], sizeof(double) * nz);
      T.qlo[a][q] = -P.max_vel, T.qhi[a][q] = P.max_vel;
    }
    for (int k = 1; k < N; ++k, ++q) {
      std::memcpy(T.cQ[a][q], T.cA[a][k], sizeof(double) * 3);
      std::memcpy(T.EQ[a][q], T.QA[a][k], sizeof(double) * nz);
      T.qlo[a][q] = alo, T.qhi[a][q] = ahi;
    }
    for (q = 0; q < T.nq; ++q) {
      bool zero = true;
      for (int z = 0; z < nz; ++z) zero &= T.EQ[a][q][z] == 0.0;
      T.qconst[a][q] = zero;
      if (!(T.qhi[a][q] > T.qlo[a][q])) return 5;
    }
    // Hessian: 2 r_u Z'Z + 2 sum_i (w_p QP_i QP_i' + w_v QV_i QV_i'), x_i tracks ref[i-1] (:870-883)
    for (int y = 0; y < nz; ++y)
      for (int z = 0; z < nz; ++z) {
        double s = 0;
        for (int c = 0; c < N; ++c) s += T.Z[a][c][y] * T.Z[a][c][z];
        s *= 2 * P.r_u;
        for (int i = 1; i <= N; ++i) {
          const double* wt = i == N ? P.r_n : P.r_x;
          s += 2 * wt[a] * T.QP[a][i][y] * T.QP[a][i][z] + 2 * wt[3 + a] * T.QV[a][i][y] * T.QV[a][i][z];
        }
        T.Hw[a][y][z] = s;
      }
    if (!detail::cholesky_inverse(T.Hw[a], nz, T.HwInv[a])) return 4;
  }
  T.nkp = 0;
  for (int k = 0; k <= N; ++k) {
    bool zero = true;
    for (int a = 0; a < 3; ++a)
      for (int z = 0; z < nz; ++z) zero &= T.QP[a][k][z] == 0.0;
    T.slot_of_kp[k] = zero ? -1 : T.nkp;
    if (!zero) T.kp_of_slot[T.nkp++] = k;
  }
  return 0;
}

}  // namespace hdsm
