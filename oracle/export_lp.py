"""Writes the reference's per-agent optimisation as a Gurobi-readable LP file (TEST INFRASTRUCTURE).

Gurobi is closed source and absent here, so parity of the optimisation against it is unpinned
(SURVEY.md section 8(c), item 5).  This exporter restates the model exactly as the reference builds it -
variables x_k (9), u_k (3) and binaries b[k][p] (agent_class.cpp:2076-2113), bounds (:2083-2097, :2179-2186),
x_0 = state_curr_ via LB = UB (:886-889), x_N[3..8] fixed to 0 (:2078-2081), dynamics equalities (:2115-2152),
objective (:2098, :870-883), one-hot rows (:939-940) and the indicator rows b[k][p] = 1 -> A x_k <= b and
A x_{k+1} <= b (:909-937) with the static rows followed by one inter-agent plane per neighbour (:1217-1234) -
so that anyone with a licence can run

    gurobi_cl Threads=1 ResultFile=agent.sol agent.lp

and compare ObjVal / the x_k values with the `obj` / `traj` written next to it (agent.expected.json).

    python -m oracle.export_lp tests/golden/config2_step8 3 /tmp/agent3     # fixture, agent index, output stem
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

from . import hdsm_oracle as o


def _lin(terms):
    """'+ 2 x - 3 y' from [(coef, name)], skipping zeros."""
    out = []
    for c, v in terms:
        if c == 0.0:
            continue
        out.append(f"{'+' if c >= 0 else '-'} {abs(c):.17g} {v}")
    return " ".join(out) if out else "0 dummy_zero"


def lp_text(p: o.Params, x0, ref, polys, planes) -> str:
    N, P = p.n_hor, min(p.poly_hor, len(polys))
    A, B = o.discrete_dynamics(p)
    X = lambda k, j: f"x_{k}_{j}"
    U = lambda k, j: f"u_{k}_{j}"
    Bv = lambda k, q: f"b_{k}_{q}"
    L = ["\\ per-agent trajectory optimisation of lis-epfl/multi_agent_pkgs (agent_class.cpp:858-1023, :2071-2153)",
         "Minimize", " obj:"]
    # objective: r_u sum u^2 + sum_{i=1..N} sum_{j<6} w_i[j] (x_i[j] - ref_{i-1}[j])^2 ; LP format: linear + [ quad ] / 2
    lin, quad, const = [], [], 0.0
    for k in range(N):
        for j in range(3):
            quad.append((2 * p.r_u, f"{U(k, j)} ^ 2"))
    for i in range(1, N + 1):
        w = p.r_n if i == N else p.r_x
        for j in range(6):
            if w[j] == 0:
                continue
            r = float(ref[i - 1][j])
            quad.append((2 * w[j], f"{X(i, j)} ^ 2"))
            lin.append((-2 * w[j] * r, X(i, j)))
            const += w[j] * r * r
    L.append("  " + _lin(lin))
    L.append("  + [ " + " ".join(f"{'+' if c >= 0 else '-'} {abs(c):.17g} {v}" for c, v in quad) + " ] / 2")
    L.append(f"  + {const:.17g} objconst")
    L.append("Subject To")
    for k in range(N):  # x_{k+1} - A x_k - B u_k = 0
        for r in range(9):
            terms = [(1.0, X(k + 1, r))] + [(-A[r, j], X(k, j)) for j in range(9)] + [(-B[r, j], U(k, j)) for j in range(3)]
            L.append(f" dyn_{k}_{r}: {_lin(terms)} = 0")
    for k in range(N):
        L.append(f" onehot_{k}: {_lin([(1.0, Bv(k, q)) for q in range(P)])} = 1")
        nk, bk = planes[k] if planes is not None else (np.zeros((0, 3)), np.zeros(0))
        for q in range(P):
            Aq, bq = polys[q]
            rowsA = np.vstack([Aq, nk]) if len(bk) else Aq
            rowsb = np.concatenate([bq, bk]) if len(bk) else bq
            for kk in (k, k + 1):
                for r in range(len(rowsb)):
                    terms = [(float(rowsA[r, j]), X(kk, j)) for j in range(3)]
                    L.append(f" poly_{k}_{q}_{kk}_{r}: {Bv(k, q)} = 1 -> {_lin(terms)} <= {float(rowsb[r]):.17g}")
    L.append("Bounds")
    L.append(" objconst = 1")
    xl, xu, ul, uu = p.x_lb(), p.x_ub(), p.u_lb(), p.u_ub()
    for k in range(N + 1):
        for j in range(9):
            if k == 0:
                L.append(f" {X(k, j)} = {float(x0[j]):.17g}")
            elif k == N and j >= 3:
                L.append(f" {X(k, j)} = 0")
            elif np.isinf(xl[j]):
                L.append(f" {X(k, j)} free")
            else:
                L.append(f" {xl[j]:.17g} <= {X(k, j)} <= {xu[j]:.17g}")
    for k in range(N):
        for j in range(3):
            L.append(f" {ul[j]:.17g} <= {U(k, j)} <= {uu[j]:.17g}")
    L.append("Binaries")
    L.append(" " + " ".join(Bv(k, q) for k in range(N) for q in range(P)))
    L.append("End")
    return "\n".join(_wrap(l) for l in L) + "\n"


def _wrap(line, width=200):
    """LP files limit the line length (510 characters in Gurobi's reader): break long rows at term boundaries."""
    if len(line) <= width:
        return line
    out, cur = [], ""
    toks = line.split(" ")
    i = 0
    while i < len(toks):
        # keep "sign coef name" triples and "name ^ 2" together: only break before a sign token
        t = toks[i]
        if t in ("+", "-") and len(cur) > width - 60:
            out.append(cur.rstrip())
            cur = "   "
        cur += t + " "
        i += 1
    out.append(cur.rstrip())
    return "\n".join(out)


def export_agent(batch, i, stem, expected=None):
    p = o.Params(**batch.params)
    lo, hi = batch.nbr_begin[i], batch.nbr_end[i]
    planes = o.time_aware_planes(p, batch.prev_self_pos[i], batch.all_pos[lo:hi], batch.all_valid[lo:hi], batch.global_id[i] - lo)
    with open(stem + ".lp", "w") as f:
        f.write(lp_text(p, batch.x0[i], batch.ref[i], batch.polys_of(i), planes))
    if expected is not None:
        with open(stem + ".expected.json", "w") as f:
            json.dump({"status": int(expected["status"][i]), "obj": float(expected["obj"][i]),
                       "traj": np.asarray(expected["traj"][i]).tolist(), "sigma": np.asarray(expected["sigma"][i]).tolist(),
                       "tolerance": "relative objective gap <= 1e-4 (Gurobi MIPGap); positions 1e-3 m"}, f)
    return stem + ".lp"


def main():
    from multi_agent_pkgs_b200.scenarios import Batch
    fixture, i, stem = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    b = Batch.load(fixture + "_in.npz")
    exp = dict(np.load(fixture + "_exp.npz")) if os.path.exists(fixture + "_exp.npz") else None
    print(export_agent(b, i, stem, exp))


if __name__ == "__main__":
    main()
