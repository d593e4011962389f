// Checks oracle/ref_shim/gurobi/gurobi_c++.h (the recording stand-in for the Gurobi C++ API) with the expression forms
// multi_agent_planner/src/agent_class.cpp uses: var - number, number * expr * expr, expr + number * expr, (a + 2 b) / 6,
// expr == expr, indicator rows, remove, attribute set / get, and the failure path of optimize().
#include <cmath>
#include <cstdio>
#include <string>
#include <vector>

#include "gurobi_c++.h"

#define CHECK(c)                                                        \
  do {                                                                  \
    if (!(c)) {                                                         \
      std::fprintf(stderr, "CHECK failed line %d: %s\n", __LINE__, #c); \
      return 1;                                                         \
    }                                                                   \
  } while (0)

int main() {
  GRBEnv env(true);
  env.set(GRB_IntParam_OutputFlag, 0);
  env.start();
  GRBModel m(env);
  // a two-step double integrator: p1 = p0 + dt v0, v1 = v0 + dt u0, cost r u0^2 + w (p1 - ref)^2
  const double dt = 0.1, r = 0.01, w = 100.0, ref = 0.7;
  GRBVar p0 = m.addVar(-GRB_INFINITY, GRB_INFINITY, 0.0, GRB_CONTINUOUS, "p0"), v0 = m.addVar(-20, 20, 0.0, GRB_CONTINUOUS, "v0");
  GRBVar p1 = m.addVar(-GRB_INFINITY, GRB_INFINITY, 0.0, GRB_CONTINUOUS, "p1"), v1 = m.addVar(0.0, 0.0, 0.0, GRB_CONTINUOUS, "v1");
  GRBVar u0 = m.addVar(-60, 60, 0.0, GRB_CONTINUOUS, "u0"), b = m.addVar(-GRB_INFINITY, GRB_INFINITY, 0.0, GRB_BINARY, "b");
  GRBQuadExpr obj;
  obj = obj + r * u0 * u0;
  GRBLinExpr k1 = GRBLinExpr(v0), k2 = GRBLinExpr(v0) + (dt / 2) * GRBLinExpr(u0);
  GRBLinExpr avg = (k1 + 2 * k2 + 2 * k2 + k1) / 6;
  GRBConstr c0 = m.addConstr(GRBLinExpr(p1, 1.0) == GRBLinExpr(p0, 1.0) + dt * avg, "dyn_p");
  GRBConstr c1 = m.addConstr(GRBLinExpr(v1, 1.0) == GRBLinExpr(v0, 1.0) + dt * GRBLinExpr(u0), "dyn_v");
  GRBQuadExpr obj_i(obj);
  obj_i += w * (p1 - ref) * (p1 - ref);
  GRBLinExpr row = -2.5;
  row = row + 1.0 * GRBLinExpr(p1) + (-0.5) * GRBLinExpr(v1);
  GRBGenConstr g = m.addGenConstrIndicator(b, 1, row, GRB_LESS_EQUAL, 0, "poly");
  GRBLinExpr sum_bin;
  sum_bin = 0;
  sum_bin = sum_bin + b;
  GRBConstr one = m.addConstr(sum_bin == 1, "one");
  m.setObjective(obj_i, GRB_MINIMIZE);
  m.set("Threads", std::to_string(1));
  m.set(GRB_DoubleParam_TimeLimit, 0.08);
  p0.set(GRB_DoubleAttr_LB, 0.25);
  p0.set(GRB_DoubleAttr_UB, 0.25);

  CHECK(m.vars.size() == 6 && m.vars[3].lb == 0.0 && m.vars[3].ub == 0.0 && m.vars[5].type == GRB_BINARY && m.vars[0].lb == 0.25 && m.vars[0].ub == 0.25);
  // dyn_p:  p1 - p0 - dt (v0 + dt/3 u0 ... ) : (k1 + 4 k2 + k1) / 6 = v0 + (4 dt / 12) u0
  const GRBModel::Lin& d = m.lin[c0.index];
  CHECK(d.sense == GRB_EQUAL && d.expr.coeff_of(p1.index) == 1.0 && d.expr.coeff_of(p0.index) == -1.0);
  CHECK(std::fabs(d.expr.coeff_of(v0.index) + dt) < 1e-15 && std::fabs(d.expr.coeff_of(u0.index) + dt * (4 * dt / 2) / 6) < 1e-15 && d.expr.constant == 0.0);
  CHECK(m.lin[c1.index].expr.coeff_of(u0.index) == -dt && m.lin[c1.index].name == "dyn_v");
  // objective: r u0^2 + w p1^2 - 2 w ref p1 + w ref^2
  CHECK(m.objective.quad_coeff(u0.index, u0.index) == r && m.objective.quad_coeff(p1.index, p1.index) == w && m.objective.quad_coeff(p1.index, u0.index) == 0.0);
  CHECK(std::fabs(m.objective.lin.coeff_of(p1.index) + 2 * w * ref) < 1e-12 && std::fabs(m.objective.lin.constant - w * ref * ref) < 1e-12);
  CHECK(m.obj_sense == GRB_MINIMIZE && m.env.str_params["Threads"] == "1" && m.env.dbl_params[GRB_DoubleParam_TimeLimit] == 0.08);
  // indicator row and the one-hot row
  const GRBModel::Ind& in = m.ind[g.index];
  CHECK(in.bin_var == b.index && in.bin_val == 1 && in.sense == GRB_LESS_EQUAL && in.rhs == 0 && in.expr.constant == -2.5 && in.expr.coeff_of(p1.index) == 1.0 &&
        in.expr.coeff_of(v1.index) == -0.5);
  CHECK(m.lin[one.index].expr.coeff_of(b.index) == 1.0 && m.lin[one.index].expr.constant == -1.0);
  // remove / re-add as the reference does every step
  m.remove(one);
  m.remove(g);
  CHECK(m.n_active_lin() == 2 && m.n_active_ind() == 0);
  // no solver: optimize throws what the reference catches; with a hook the solution is readable through the variables
  bool threw = false;
  try {
    m.optimize();
  } catch (GRBException e) {
    threw = e.getErrorCode() != 0 && !e.getMessage().empty();
  }
  CHECK(threw && m.optimize_calls == 1);
  GRBModel::solver_hook() = [](GRBModel& mm) { for (size_t i = 0; i < mm.vars.size(); ++i) mm.vars[i].x = 10.0 + i; mm.status = GRB_OPTIMAL; };
  m.optimize();
  CHECK(u0.get(GRB_DoubleAttr_X) == 14.0 && b.get(GRB_DoubleAttr_X) == 15.0 && m.get(GRB_IntAttr_Status) == GRB_OPTIMAL);
  GRBModel::solver_hook() = nullptr;
  std::printf("gurobi stand-in ok\n");
  return 0;
}
