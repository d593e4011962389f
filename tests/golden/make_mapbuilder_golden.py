"""Regenerates tests/golden/mapbuilder_ref.npz from the REFERENCE ITSELF.

Needs oracle/_ref/libref_map.so, i.e. the reference's own mapping_util/src/map_builder.cpp (with path_tools.cpp,
raycast.cpp, voxel_grid.cpp) compiled unmodified on the stand-in ROS / Eigen headers of oracle/ref_shim/
(`make -C oracle ref`, only possible where /root/reference is mounted).  Every case is one agent's chain of three
MapBuilder::EnvironmentVoxelGridCallback calls (map_builder.cpp:80-240) in a small random environment: the node's
voxel_grid_curr_ after each call (crop, RaycastAndClear, MergeVoxelGrids), its origin, and the grid it published
(after SetUncertainToUnknown, InflateObstacles, CreatePotentialField).  The committed file lets the checker and GPU tests
pin themselves to the reference on machines where it is absent.

    python tests/golden/make_mapbuilder_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import sensing as S  # noqa: E402

CASES = [  # voxel, range, free_grid, fov (None = 360 degrees), inflation, potential, power
    (0.3, (6.0, 5.1, 3.0), False, None, 0.3, 1.5, 4.0),
    (0.3, (6.0, 5.1, 3.0), False, (1.2, 0.8), 0.3, 1.5, 4.0),
    (0.3, (6.0, 5.1, 3.0), True, None, 0.3, 1.5, 4.0),
    (0.2, (4.0, 4.0, 2.0), False, None, 0.2, 0.6, 2.0),
    (0.5, (10.0, 8.0, 0.5), False, None, 0.5, 1.0, 1.0),
    (0.3, (6.0, 5.1, 3.0), False, (1.57, 1.57), 0.0, 0.9, 4.0),
]


def environment(rng, vox):
    ex, ey, ez = rng.integers(30, 60), rng.integers(30, 60), rng.integers(6, 16)
    env = np.zeros((ez, ey, ex), np.int8)
    for _ in range(rng.integers(5, 40)):
        env[:, rng.integers(0, ey), rng.integers(0, ex)] = 100
    for _ in range(rng.integers(0, 6)):
        x, y, z = rng.integers(0, ex), rng.integers(0, ey), rng.integers(0, ez)
        s = rng.integers(1, 5, 3)
        env[z:z + s[2], y:y + s[1], x:x + s[0]] = 100
    if rng.random() < 0.5:
        env[0] = 100
    origin = np.round(rng.uniform(-3, 3, 3) / vox) * vox
    return env, origin


def main():
    rng = np.random.default_rng(20261017)
    out = {}
    k = 0
    for vox, rng3, free, fov, infl, pot, pw in CASES:
        for rep in range(3):
            env, org = environment(rng, vox)
            hi = org + np.array(env.shape[::-1]) * vox
            pos = rng.uniform(org + 0.5, hi - 0.5)
            yaw = rng.uniform(-np.pi, np.pi)
            rot = np.array([[np.cos(yaw), -np.sin(yaw), 0], [np.sin(yaw), np.cos(yaw), 0], [0, 0, 1.0]])
            old_g = old_o = None
            steps = []
            for step in range(3):
                cur, o, pub = S.ref_update(env, org, pos, vox, rng3, free_grid=free, rot=rot if fov else None, fov=fov, old_grid=old_g,
                                           old_origin=old_o, inflation=infl, potential=pot, power=pw)
                steps.append((pos.copy(), cur, o, pub))
                old_g, old_o = cur, o
                pos = pos + rng.uniform(-0.8, 0.8, 3) * [1, 1, 0.2]
            out[f"c{k}_prm"] = np.array([vox, *rng3, float(free), float(fov is not None), *(fov or (0, 0)), infl, pot, pw])
            out[f"c{k}_env"], out[f"c{k}_org"], out[f"c{k}_rot"] = env, org, rot
            out[f"c{k}_pos"] = np.stack([s[0] for s in steps])
            out[f"c{k}_cur"] = np.stack([s[1] for s in steps])
            out[f"c{k}_origin"] = np.stack([s[2] for s in steps])
            out[f"c{k}_pub"] = np.stack([s[3] for s in steps])
            k += 1
    out["n_cases"] = np.array(k)
    path = os.path.join(ROOT, "tests", "golden", "mapbuilder_ref.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {k} cases x 3 updates to {path} ({os.path.getsize(path) / 1024:.0f} KB)")


if __name__ == "__main__":
    main()
