// TEST INFRASTRUCTURE - stand-in for the Eigen-based geometry headers of the reference.
//
// The reference's corridor generator (convex_decomp_util/src/convex_decomp.cpp) only needs small
// fixed-size integer / double vectors and a container of (point, normal) pairs from
// decomp_geometry/polyhedron.h and decomp_basis/data_type.h, which in the reference are Eigen aliases.
// Eigen is not installed in this image, so this header supplies just those types with the same names
// and the handful of operations the generator uses (element access with () and [], +, -, unary -,
// dot).  With it on the include path the reference's own convex_decomp.cpp compiles UNMODIFIED from
// /root/reference into oracle/_ref/ (see oracle/Makefile target `ref`), which is what pins the C
// restatement in oracle/corridor_oracle.c and the CUDA kernel.  Nothing of the product includes this.
#ifndef HDSM_REF_SHIM_POLYHEDRON_H_
#define HDSM_REF_SHIM_POLYHEDRON_H_

#include <math.h>
#include <stdint.h>
#include <stdio.h>

#include <utility>
#include <vector>

#include <Eigen/Dense>  // the stand-in next to this directory: in the reference Vec3f / Vec3i ARE Eigen::Vector3d / Vector3i

typedef double decimal_t;

template <class S, int N>
struct ShimVec {
  S v[N];
  ShimVec() {}
  ShimVec(S a, S b) { static_assert(N == 2, "size"); v[0] = a, v[1] = b; }
  ShimVec(S a, S b, S c) { static_assert(N == 3, "size"); v[0] = a, v[1] = b, v[2] = c; }
  ShimVec(S a, S b, S c, S d) { static_assert(N == 4, "size"); v[0] = a, v[1] = b, v[2] = c, v[3] = d; }
  S& operator()(int i) { return v[i]; }
  const S& operator()(int i) const { return v[i]; }
  S& operator[](int i) { return v[i]; }
  const S& operator[](int i) const { return v[i]; }
  ShimVec operator+(const ShimVec& o) const { ShimVec r; for (int i = 0; i < N; ++i) r.v[i] = v[i] + o.v[i]; return r; }
  ShimVec operator-(const ShimVec& o) const { ShimVec r; for (int i = 0; i < N; ++i) r.v[i] = v[i] - o.v[i]; return r; }
  ShimVec operator-() const { ShimVec r; for (int i = 0; i < N; ++i) r.v[i] = -v[i]; return r; }
  S dot(const ShimVec& o) const { S s = v[0] * o.v[0]; for (int i = 1; i < N; ++i) s += v[i] * o.v[i]; return s; }
};

// size 3 is the Eigen stand-in's vector, so that code mixing Vec3f and Eigen::Vector3d (multi_agent_planner) sees one type
template <class S, int N> struct ShimVecSel { typedef ShimVec<S, N> type; };
template <> struct ShimVecSel<int, 3> { typedef Eigen::Vector3i type; };
template <> struct ShimVecSel<decimal_t, 3> { typedef Eigen::Vector3d type; };
template <int N> using Veci = typename ShimVecSel<int, N>::type;
template <int N> using Vecf = typename ShimVecSel<decimal_t, N>::type;
typedef Veci<2> Vec2i;
typedef Veci<3> Vec3i;
typedef Vecf<2> Vec2f;
typedef Vecf<3> Vec3f;
template <class T> using vec_E = std::vector<T>;

template <int Dim>
struct Hyperplane {
  Hyperplane() {}
  Hyperplane(const Vecf<Dim>& p, const Vecf<Dim>& n) : p_(p), n_(n) {}
  Vecf<Dim> p_, n_;
};
typedef Hyperplane<3> Hyperplane3D;

template <int Dim>
struct Polyhedron {
  Polyhedron() {}
  Polyhedron(const vec_E<Hyperplane<Dim>>& vs) : vs_(vs) {}
  vec_E<std::pair<Vecf<Dim>, Vecf<Dim>>> cal_normals() const {
    vec_E<std::pair<Vecf<Dim>, Vecf<Dim>>> ns(vs_.size());
    for (size_t i = 0; i < vs_.size(); i++) ns[i] = std::make_pair(vs_[i].p_, vs_[i].n_);
    return ns;
  }
  vec_E<Hyperplane<Dim>> hyperplanes() const { return vs_; }
  vec_E<Hyperplane<Dim>> vs_;
};
typedef Polyhedron<3> Polyhedron3D;

// dynamic-row matrix / vector and the linear-constraint container (decomp_basis/data_type.h, polyhedron.h:98-147), as far as
// multi_agent_planner/src/agent_class.cpp uses them: element access, rows / cols / size, row assignment, conservativeResize
template <int N>
struct MatDNf {
  int r;
  std::vector<decimal_t> d;  // row-major r x N
  struct Row {
    MatDNf* M; int i;
    Row& operator=(const Vecf<N>& v) { for (int c = 0; c < N; ++c) M->d[(size_t)i * N + c] = v(c); return *this; }
  };
  MatDNf() : r(0) {}
  MatDNf(int rows, int cols) : r(rows), d((size_t)rows * N, 0.0) { (void)cols; }
  int rows() const { return r; }
  int cols() const { return N; }
  decimal_t& operator()(int i, int j) { return d[(size_t)i * N + j]; }
  decimal_t operator()(int i, int j) const { return d[(size_t)i * N + j]; }
  Row row(int i) { Row k = {this, i}; return k; }
  void conservativeResize(int rows, int cols) { (void)cols; r = rows; d.resize((size_t)rows * N, 0.0); }
};
struct VecDf {
  std::vector<decimal_t> d;
  VecDf() {}
  explicit VecDf(int n) : d(n, 0.0) {}
  int size() const { return (int)d.size(); }
  int rows() const { return (int)d.size(); }
  decimal_t& operator()(int i) { return d[i]; }
  decimal_t operator()(int i) const { return d[i]; }
  decimal_t& operator[](int i) { return d[i]; }
  decimal_t operator[](int i) const { return d[i]; }
  void conservativeResize(int n) { d.resize(n, 0.0); }
};
template <int Dim>
struct LinearConstraint {
  LinearConstraint() {}
  LinearConstraint(const MatDNf<Dim>& A, const VecDf& b) : A_(A), b_(b) {}
  bool inside(const Vecf<Dim>& pt) const {
    for (int i = 0; i < A_.rows(); ++i) {
      decimal_t s = -b_(i);
      for (int c = 0; c < Dim; ++c) s += A_(i, c) * pt(c);
      if (s > 0) return false;
    }
    return true;
  }
  MatDNf<Dim> A() const { return A_; }
  VecDf b() const { return b_; }
  MatDNf<Dim> A_;
  VecDf b_;
};
typedef LinearConstraint<3> LinearConstraint3D;

#endif
