// TEST INFRASTRUCTURE - RECORDING stand-in for the Gurobi C++ API (gurobi_c++.h of Gurobi 10, which the reference links
// and which is closed source, licence-gated and absent here).  It implements the part of the API that
// multi_agent_planner/src/agent_class.cpp uses (:858-1023, :1071-1084, :2063-2167): variables, linear and quadratic
// expressions with their operators, linear constraints, indicator constraints, objective, parameters, attribute get /
// set - and instead of solving it RECORDS the model, so that the model the reference's own code builds can be read back
// and compared with what the CUDA library is given (objective matrix, dynamics rows, indicator rows).  optimize() calls
// GRBModel::solver_hook when a test installs one (it may fill GRBModel::vars[i].x) and throws GRBException otherwise,
// which the reference treats as a failed optimisation (:988-995).  My own code; nothing of the product includes this.
// Groundwork for compiling agent_class.cpp unmodified the way map_builder.cpp already is.
#ifndef HDSM_REF_SHIM_GUROBI_CPP_H_
#define HDSM_REF_SHIM_GUROBI_CPP_H_
#include <functional>
#include <map>
#include <string>
#include <utility>
#include <vector>

#define GRB_INFINITY 1e100
#define GRB_CONTINUOUS 'C'
#define GRB_BINARY 'B'
#define GRB_INTEGER 'I'
#define GRB_LESS_EQUAL '<'
#define GRB_GREATER_EQUAL '>'
#define GRB_EQUAL '='
#define GRB_MINIMIZE 1
#define GRB_MAXIMIZE (-1)
#define GRB_OPTIMAL 2
#define GRB_INFEASIBLE 3
#define GRB_TIME_LIMIT 9

enum GRB_DoubleAttr { GRB_DoubleAttr_X, GRB_DoubleAttr_LB, GRB_DoubleAttr_UB, GRB_DoubleAttr_Obj, GRB_DoubleAttr_ObjVal, GRB_DoubleAttr_Start };
enum GRB_IntAttr { GRB_IntAttr_Status, GRB_IntAttr_NumVars, GRB_IntAttr_NumConstrs, GRB_IntAttr_SolCount };
enum GRB_IntParam { GRB_IntParam_OutputFlag, GRB_IntParam_Threads, GRB_IntParam_Presolve };
enum GRB_DoubleParam { GRB_DoubleParam_TimeLimit, GRB_DoubleParam_MIPGap, GRB_DoubleParam_FeasibilityTol, GRB_DoubleParam_OptimalityTol };

class GRBException {
 public:
  GRBException(const std::string& m = "", int c = 0) : msg_(m), code_(c) {}
  std::string getMessage() const { return msg_; }
  int getErrorCode() const { return code_; }
 private:
  std::string msg_;
  int code_;
};

class GRBModel;

class GRBVar {
 public:
  GRBVar() : model(nullptr), index(-1) {}
  GRBVar(GRBModel* m, int i) : model(m), index(i) {}
  double get(GRB_DoubleAttr a) const;
  void set(GRB_DoubleAttr a, double v);
  GRBModel* model;
  int index;
};

class GRBLinExpr {
 public:
  GRBLinExpr(double c = 0.0) : constant(c) {}
  GRBLinExpr(GRBVar v, double coeff = 1.0) : constant(0.0) { coef[v.index] = coeff; }
  GRBLinExpr& operator+=(const GRBLinExpr& o) { for (auto& kv : o.coef) coef[kv.first] += kv.second; constant += o.constant; return *this; }
  GRBLinExpr& operator-=(const GRBLinExpr& o) { for (auto& kv : o.coef) coef[kv.first] -= kv.second; constant -= o.constant; return *this; }
  GRBLinExpr& operator*=(double s) { for (auto& kv : coef) kv.second *= s; constant *= s; return *this; }
  GRBLinExpr operator-() const { GRBLinExpr r(*this); r *= -1.0; return r; }
  double coeff_of(int var) const { auto it = coef.find(var); return it == coef.end() ? 0.0 : it->second; }
  std::map<int, double> coef;  // variable index -> coefficient
  double constant;
};
inline GRBLinExpr operator+(GRBLinExpr a, const GRBLinExpr& b) { a += b; return a; }
inline GRBLinExpr operator-(GRBLinExpr a, const GRBLinExpr& b) { a -= b; return a; }
inline GRBLinExpr operator*(double s, GRBLinExpr a) { a *= s; return a; }
inline GRBLinExpr operator*(GRBLinExpr a, double s) { a *= s; return a; }
inline GRBLinExpr operator/(GRBLinExpr a, double s) { a *= 1.0 / s; return a; }
inline GRBLinExpr operator+(GRBVar a, GRBVar b) { return GRBLinExpr(a) + GRBLinExpr(b); }
inline GRBLinExpr operator-(GRBVar a, GRBVar b) { return GRBLinExpr(a) - GRBLinExpr(b); }
inline GRBLinExpr operator+(GRBVar a, double c) { return GRBLinExpr(a) + GRBLinExpr(c); }
inline GRBLinExpr operator-(GRBVar a, double c) { return GRBLinExpr(a) - GRBLinExpr(c); }
inline GRBLinExpr operator*(double s, GRBVar v) { return GRBLinExpr(v, s); }
inline GRBLinExpr operator*(GRBVar v, double s) { return GRBLinExpr(v, s); }

class GRBQuadExpr {
 public:
  GRBQuadExpr(double c = 0.0) : lin(c) {}
  GRBQuadExpr(GRBVar v) : lin(v) {}
  GRBQuadExpr(const GRBLinExpr& l) : lin(l) {}
  GRBQuadExpr& operator+=(const GRBQuadExpr& o) { lin += o.lin; for (auto& kv : o.quad) quad[kv.first] += kv.second; return *this; }
  GRBQuadExpr& operator-=(const GRBQuadExpr& o) { lin -= o.lin; for (auto& kv : o.quad) quad[kv.first] -= kv.second; return *this; }
  GRBQuadExpr& operator*=(double s) { lin *= s; for (auto& kv : quad) kv.second *= s; return *this; }
  double quad_coeff(int i, int j) const { auto it = quad.find(i <= j ? std::make_pair(i, j) : std::make_pair(j, i)); return it == quad.end() ? 0.0 : it->second; }
  GRBLinExpr lin;
  std::map<std::pair<int, int>, double> quad;  // (i <= j) -> coefficient of x_i x_j
};
inline GRBQuadExpr operator*(const GRBLinExpr& a, const GRBLinExpr& b) {
  GRBQuadExpr q(a.constant * b.constant);
  for (auto& ka : a.coef) q.lin.coef[ka.first] += ka.second * b.constant;
  for (auto& kb : b.coef) q.lin.coef[kb.first] += kb.second * a.constant;
  for (auto& ka : a.coef)
    for (auto& kb : b.coef) {
      const int i = ka.first, j = kb.first;
      q.quad[i <= j ? std::make_pair(i, j) : std::make_pair(j, i)] += ka.second * kb.second;
    }
  return q;
}
inline GRBQuadExpr operator*(const GRBLinExpr& a, GRBVar b) { return a * GRBLinExpr(b); }
inline GRBQuadExpr operator*(GRBVar a, const GRBLinExpr& b) { return GRBLinExpr(a) * b; }
inline GRBQuadExpr operator*(GRBVar a, GRBVar b) { return GRBLinExpr(a) * GRBLinExpr(b); }
inline GRBQuadExpr operator+(GRBQuadExpr a, const GRBQuadExpr& b) { a += b; return a; }
inline GRBQuadExpr operator-(GRBQuadExpr a, const GRBQuadExpr& b) { a -= b; return a; }
inline GRBQuadExpr operator*(double s, GRBQuadExpr a) { a *= s; return a; }
inline GRBQuadExpr operator*(GRBQuadExpr a, double s) { a *= s; return a; }

struct GRBTempConstr { GRBLinExpr expr; char sense; };  // expr (sense) 0
inline GRBTempConstr operator==(const GRBLinExpr& a, const GRBLinExpr& b) { return GRBTempConstr{a - b, GRB_EQUAL}; }
inline GRBTempConstr operator<=(const GRBLinExpr& a, const GRBLinExpr& b) { return GRBTempConstr{a - b, GRB_LESS_EQUAL}; }
inline GRBTempConstr operator>=(const GRBLinExpr& a, const GRBLinExpr& b) { return GRBTempConstr{a - b, GRB_GREATER_EQUAL}; }

class GRBConstr { public: GRBConstr(int i = -1) : index(i) {} int index; };
class GRBGenConstr { public: GRBGenConstr(int i = -1) : index(i) {} int index; };

class GRBEnv {
 public:
  explicit GRBEnv(bool empty = false) : started(!empty) {}
  void set(GRB_IntParam p, int v) { int_params[p] = v; }
  void set(GRB_DoubleParam p, double v) { dbl_params[p] = v; }
  void set(const std::string& k, const std::string& v) { str_params[k] = v; }
  void start() { started = true; }
  bool started;
  std::map<int, int> int_params;
  std::map<int, double> dbl_params;
  std::map<std::string, std::string> str_params;
};

class GRBModel {
 public:
  struct Var { double lb, ub, obj; char type; std::string name; double x; };
  struct Lin { GRBLinExpr expr; char sense; std::string name; bool removed; };                               // expr (sense) 0
  struct Ind { int bin_var, bin_val; GRBLinExpr expr; char sense; double rhs; std::string name; bool removed; };  // bin == val -> expr (sense) rhs
  explicit GRBModel(const GRBEnv& e) : env(e), obj_sense(GRB_MINIMIZE), status(0), optimize_calls(0) {}
  GRBVar addVar(double lb, double ub, double obj, char type, const std::string& name = "") {
    vars.push_back(Var{lb, ub, obj, type, name, 0.0});
    return GRBVar(this, (int)vars.size() - 1);
  }
  GRBConstr addConstr(const GRBTempConstr& t, const std::string& name = "") {
    lin.push_back(Lin{t.expr, t.sense, name, false});
    return GRBConstr((int)lin.size() - 1);
  }
  GRBGenConstr addGenConstrIndicator(GRBVar bin, int val, const GRBLinExpr& e, char sense, double rhs, const std::string& name = "") {
    ind.push_back(Ind{bin.index, val, e, sense, rhs, name, false});
    return GRBGenConstr((int)ind.size() - 1);
  }
  void remove(GRBConstr c) { lin.at(c.index).removed = true; }
  void remove(GRBGenConstr c) { ind.at(c.index).removed = true; }
  void setObjective(const GRBQuadExpr& q, int sense = GRB_MINIMIZE) { objective = q; obj_sense = sense; }
  void set(GRB_IntParam p, int v) { env.set(p, v); }
  void set(GRB_DoubleParam p, double v) { env.set(p, v); }
  void set(const std::string& k, const std::string& v) { env.set(k, v); }
  void update() {}
  int get(GRB_IntAttr a) const { return a == GRB_IntAttr_Status ? status : a == GRB_IntAttr_NumVars ? (int)vars.size() : a == GRB_IntAttr_NumConstrs ? n_active_lin() : 0; }
  int n_active_lin() const { int n = 0; for (auto& c : lin) n += !c.removed; return n; }
  int n_active_ind() const { int n = 0; for (auto& c : ind) n += !c.removed; return n; }
  void optimize() {
    ++optimize_calls;
    if (solver_hook()) { solver_hook()(*this); return; }
    throw GRBException("recording stand-in: no solver installed", 10009);
  }
  static std::function<void(GRBModel&)>& solver_hook() { static std::function<void(GRBModel&)> h; return h; }
  GRBEnv env;
  std::vector<Var> vars;
  std::vector<Lin> lin;
  std::vector<Ind> ind;
  GRBQuadExpr objective;
  int obj_sense, status, optimize_calls;
};

inline double GRBVar::get(GRB_DoubleAttr a) const {
  const GRBModel::Var& v = model->vars.at(index);
  return a == GRB_DoubleAttr_X ? v.x : a == GRB_DoubleAttr_LB ? v.lb : a == GRB_DoubleAttr_UB ? v.ub : v.obj;
}
inline void GRBVar::set(GRB_DoubleAttr a, double val) {
  GRBModel::Var& v = model->vars.at(index);
  if (a == GRB_DoubleAttr_LB) v.lb = val;
  else if (a == GRB_DoubleAttr_UB) v.ub = val;
  else if (a == GRB_DoubleAttr_Obj) v.obj = val;
  else if (a == GRB_DoubleAttr_X || a == GRB_DoubleAttr_Start) v.x = val;
}
#endif
