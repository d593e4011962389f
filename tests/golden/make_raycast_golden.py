"""Regenerates tests/golden/raycast_ref.npz from the REFERENCE ITSELF (voxel_grid_util::Raycast, raycast.cpp:21-186,
compiled unmodified with voxel_grid.cpp into oracle/_ref/libref_voxel.so by `make -C oracle ref`).

    python tests/golden/make_raycast_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import reftraj as rt  # noqa: E402


def random_case(rng, t):
    dx, dy, dz = (int(v) for v in rng.integers(4, 28, 3))
    g = np.zeros((dz, dy, dx), np.int8)
    g[rng.random(g.shape) < rng.choice([0.0, 0.01, 0.05, 0.3])] = 100
    g[(rng.random(g.shape) < 0.05) & (g == 0)] = rng.integers(1, 99)
    if t % 7 == 0:
        g[rng.random(g.shape) < 0.02] = -1
    lo, hi = -2.0, np.array([dx, dy, dz]) + 2.0
    s, e = rng.uniform(lo, hi), rng.uniform(lo, hi)
    k = t % 6
    if k == 1:
        e[0] = s[0]                                             # axis parallel
    if k == 2:
        e[:2] = s[:2]
    if k == 3:
        s, e = np.floor(s) + 0.5, np.floor(e) + 0.5             # voxel centres
    if k == 4:
        s, e = np.floor(s), np.floor(e) + rng.choice([0.0, 0.25])  # on voxel boundaries
    if k == 5:
        e = s + rng.uniform(-0.4, 0.4, 3)                       # inside one voxel or its neighbour
    md = float(np.linalg.norm(s - e)) * rng.choice([1.0, 1.0, 0.5, 2.0])
    return g, s, e, md


def main(n_cases=160, seed=7):
    assert rt.have_ref(), "the reference checkout is needed to regenerate this fixture (make -C oracle ref)"
    rng = np.random.default_rng(seed)
    out = {"n_cases": np.array(n_cases)}
    for k in range(n_cases):
        g, s, e, md = random_case(rng, k)
        vis, col, n = rt.ref_raycast(g, s, e, md)
        out[f"grid{k}"], out[f"ray{k}"], out[f"vis{k}"], out[f"col{k}"] = g, np.array([*s, *e, md]), vis, col
    path = os.path.join(ROOT, "tests", "golden", "raycast_ref.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
