"""The Gurobi-readable export of the reference model (oracle/export_lp.py): the golden optimum must satisfy
every row of the written file (dynamics, one-hot, the indicator rows of the chosen polytopes, bounds) and
reproduce the objective - evaluated by a small LP-text parser, so the file itself is what is checked."""
import re

import numpy as np

from conftest import load_golden
from oracle import export_lp, hdsm_oracle as o


def _terms(expr):
    return [((-1.0 if sg == "-" else 1.0) * float(c), v) for sg, c, v in re.findall(r"([+-])\s*([0-9.eE+-]+)\s+([a-z_0-9]+)", expr)]


def test_golden_optimum_satisfies_the_exported_model(tmp_path):
    b, exp = load_golden("config2_step8")
    i = int(np.flatnonzero(exp["status"] == 0)[0])
    stem = str(tmp_path / "agent")
    export_lp.export_agent(b, i, stem, exp)
    text = open(stem + ".lp").read()
    assert max(len(l) for l in text.split("\n")) <= 510
    text = re.sub(r"\n\s{3,}", " ", text)  # undo the line wrapping
    p = o.Params(**b.params)
    N = p.n_hor
    val = {"objconst": 1.0}
    for k in range(N + 1):
        for j in range(9):
            val[f"x_{k}_{j}"] = exp["traj"][i][k, j]
    for k in range(N):
        for j in range(3):
            val[f"u_{k}_{j}"] = exp["ctrl"][i][k, j]
        for q in range(p.poly_hor):
            val[f"b_{k}_{q}"] = 1.0 if exp["sigma"][i][k] == q else 0.0
    n_dyn = n_ind = 0
    for line in text.split("\n"):
        m = re.match(r"\s*(dyn|onehot|poly)_[0-9_]+:\s*(.*)", line)
        if not m:
            continue
        body = m.group(2)
        if m.group(1) == "poly":
            guard, body = body.split("->")
            n_ind += 1
            if val[guard.split("=")[0].strip()] != 1.0:
                continue
        op = "<=" if "<=" in body else "="
        lhs, rhs = body.split(op)
        v = sum(c * val[name] for c, name in _terms(" " + lhs))
        if op == "=":
            n_dyn += m.group(1) == "dyn"
            assert abs(v - float(rhs)) <= 1e-7, line[:80]
        else:
            assert v <= float(rhs) + 1e-6, line[:80]
    assert n_dyn == 9 * N
    polys = b.polys_of(i)
    n_nb = int(b.all_valid[b.nbr_begin[i]:b.nbr_end[i]].sum()) - 1
    assert n_ind == 2 * N * sum(len(bq) + n_nb for _, bq in polys)
    # objective: linear + quadratic / 2 + constant
    obj_block = text[text.index("obj:") + 4:text.index("Subject To")]
    lin_part, rest = obj_block.split("+ [")
    quad_part, tail = rest.split("] / 2")
    lin = sum(c * val[name] for c, name in _terms(lin_part))
    quad = sum((-1.0 if sg == "-" else 1.0) * float(c) * val[v] ** 2 for sg, c, v in re.findall(r"([+-])\s*([0-9.eE+-]+)\s+([a-z_0-9]+) \^ 2", quad_part))
    const = sum(c * val[name] for c, name in _terms(tail))
    assert abs(lin + quad / 2 + const - exp["obj"][i]) <= 1e-6 * max(1.0, abs(exp["obj"][i]))
