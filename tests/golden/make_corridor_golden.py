"""Regenerates tests/golden/corridor_ref.npz from the REFERENCE ITSELF.

Needs oracle/_ref/libref_corridor.so, i.e. the reference's own convex_decomp_util/src/convex_decomp.cpp
compiled unmodified (`make -C oracle ref`, only possible where /root/reference is mounted).  Every case is
a random voxel grid, a seed and the parameters of one GetPolyOcta3D call (convex_decomp.cpp:5-376) with the
hyperplanes (point, normal) and the marked grid the reference returned.  The committed file lets the
oracle and GPU tests pin themselves to the reference on machines where it is absent.

    python tests/golden/make_corridor_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import corridor as oc  # noqa: E402


def random_grid(rng, kind, dx, dy, dz):
    g = np.zeros((dz, dy, dx), np.int8)
    if kind == 0:    # forest of inflated columns
        for _ in range(rng.integers(1, max(2, dx * dy // 40))):
            x, y = rng.integers(0, dx), rng.integers(0, dy)
            w = rng.integers(1, 3)
            g[:, max(0, y - w):y + w + 1, max(0, x - w):x + w + 1] = 100
    elif kind == 1:  # boxes
        for _ in range(rng.integers(1, 30)):
            x, y, z = rng.integers(0, dx), rng.integers(0, dy), rng.integers(0, dz)
            s = rng.integers(1, 5, 3)
            g[z:z + s[2], y:y + s[1], x:x + s[0]] = 100
    elif kind == 2:  # salt noise plus potential-field values (free for the decomposition)
        g[rng.random(g.shape) < rng.uniform(0.001, 0.05)] = 100
        g[(rng.random(g.shape) < 0.1) & (g == 0)] = 37
    else:            # slanted wall
        X, Y = np.meshgrid(np.arange(dx), np.arange(dy))
        a, b = rng.uniform(-2, 2), rng.uniform(0, dy)
        g[:, (Y > a * X + b) & (Y < a * X + b + 3)] = 100
    return g


def main(n_cases=112, n_original=64, seed=2026):
    """Cases 0..63: GetPolyOcta3D; cases 64..111: GetPolyOcta3DNew (convex_decomp.cpp:590-1162), two thirds of
    them with the seed squeezed between two occupied voxels as in agent_class.cpp:1385-1395."""
    assert oc.build_ref() and oc.have_ref(), "the reference checkout is needed to regenerate this fixture"
    rng = np.random.default_rng(seed)
    out = {}
    k = 0
    while k < n_cases:
        use_new = k >= n_original
        dx, dy, dz = (int(v) for v in (rng.integers(12, 44), rng.integers(12, 44), rng.integers(8, 20)))
        g = random_grid(rng, k % 4, dx, dy, dz)
        free = np.argwhere(g < 100)
        if len(free) == 0:
            continue
        z, y, x = (int(v) for v in free[rng.integers(len(free))])
        if use_new and k % 3 != 0:
            ax = int(rng.integers(3))
            for sgn in (-1, 1):
                c = [x, y, z]
                c[ax] += sgn
                if 0 <= c[0] < dx and 0 <= c[1] < dy and 0 <= c[2] < dz:
                    g[c[2], c[1], c[0]] = 100
        n_it = int(rng.choice([6, 17, 42, 42, 60, 90]))
        res = float(rng.choice([0.3, 0.2, 0.25]))
        conv = -int(rng.integers(1, 5))
        origin = np.round(rng.uniform(-20, 20, 3) / res) * res
        pts, nrm, marked = oc.ref_poly(g, (x, y, z), n_it, res, conv, origin, use_new=use_new)
        out[f"grid{k}"] = g
        out[f"call{k}"] = np.array([x, y, z, n_it, conv, int(use_new)], np.int32)
        out[f"fp{k}"] = np.array([res, *origin])
        out[f"pts{k}"] = pts
        out[f"nrm{k}"] = nrm
        out[f"marks{k}"] = np.flatnonzero(marked.ravel() == conv).astype(np.int32)
        k += 1
    out["n_cases"] = np.array(n_cases)
    path = os.path.join(ROOT, "tests", "golden", "corridor_ref.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
