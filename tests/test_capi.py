"""The C-ABI library: it builds, loads, exports every symbol include/hdsm.h declares, and its
struct layouts match the ctypes mirrors.  No compute calls here (those are the gpu tests)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, has_gpu
from multi_agent_pkgs_b200 import _build, _lib, scenarios as sc


def test_library_builds_and_loads():
    path = _build.build_lib()
    assert os.path.exists(path) and path.startswith(os.path.join(ROOT, "multi_agent_pkgs_b200"))
    L = _lib.load()
    assert L.hdsm_version() == 100


def test_exports_match_header():
    hdr = open(os.path.join(ROOT, "include", "hdsm.h")).read()
    declared = sorted(set(re.findall(r"\b(hdsm_[a-z_]+)\s*\(", hdr)))
    assert declared == sorted(_lib.EXPORTS)
    L = _lib.load()
    for sym in declared:
        assert hasattr(L, sym), sym


def test_header_cites_reference_lines():
    hdr = open(os.path.join(ROOT, "include", "hdsm.h")).read()
    for cite in (":2071-2153", ":1086-1215", ":858-1023", ":645-677"):
        assert cite in hdr


def test_struct_layouts():
    assert C.sizeof(_lib.HdsmResult) == 32 and _lib.RESULT_DTYPE.itemsize == 32
    assert C.sizeof(_lib.HdsmParams) == 8 * 4 + 8 * (1 + 3 + 1 + 6 + 6 + 6 + 3 + 1) + 2 * 4
    p = _lib.make_params(sc.agile_params())
    assert (p.n_hor, p.poly_hor, p.max_rows_per_poly) == (10, 4, 18) and p.max_jerk == 60.0


def test_create_rejects_bad_arguments():
    L = _lib.load()
    h = C.c_void_p()
    bad = _lib.make_params(sc.agile_params())
    bad.n_hor = 40
    assert L.hdsm_create(C.byref(bad), 1, 1, 0, C.byref(h)) == -1 and not h.value
    assert L.hdsm_create(None, 1, 1, 0, C.byref(h)) == -1


@pytest.mark.skipif(has_gpu(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback_without_device():
    """The product path must fail loudly when there is no CUDA device."""
    from multi_agent_pkgs_b200.planner import HdsmError, TrajectoryPlanner
    L = _lib.load()
    h = C.c_void_p()
    ok = _lib.make_params(sc.agile_params())
    assert L.hdsm_create(C.byref(ok), 4, 10, 0, C.byref(h)) == -2  # HDSM_ERR_CUDA
    with pytest.raises(HdsmError):
        TrajectoryPlanner(sc.agile_params(), 4, 10)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "multi_agent_pkgs_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("# oracle-free", ""), f


def test_corridor_structs_and_argument_checks():
    """hdsm_corridor_params layout, and hdsm_corridor_create rejecting what the kernel cannot do (window of
    32^3 voxels: n_it_decomp <= 90; row stride 18..32; poly_hor 1..8)."""
    from multi_agent_pkgs_b200.corridor import HdsmCorridorParams
    assert C.sizeof(HdsmCorridorParams) == 6 * 4 + 8
    L = _lib.load()
    L.hdsm_corridor_create.restype = C.c_int
    h = C.c_void_p()

    def create(poly_hor=4, n_it=42, rmax=18, n_traj=11, max_path=16, new=0, voxel=0.3, agents=4, grids=4, stride=1000):
        p = HdsmCorridorParams(poly_hor, n_it, rmax, n_traj, max_path, new, voxel)
        return L.hdsm_corridor_create(C.byref(p), C.c_int(agents), C.c_int(grids), C.c_size_t(stride), C.c_int(0), C.byref(h))

    for bad in (dict(n_it=91), dict(rmax=17), dict(rmax=33), dict(poly_hor=0), dict(poly_hor=9), dict(voxel=0.0),
                dict(max_path=0), dict(agents=0), dict(grids=0), dict(stride=0)):
        assert create(**bad) == -1, bad  # HDSM_ERR_INVALID
    assert L.hdsm_corridor_create(None, 1, 1, C.c_size_t(8), 0, C.byref(h)) == -1
    if not has_gpu():
        assert create() == -2  # HDSM_ERR_CUDA: no CPU fallback
        from multi_agent_pkgs_b200.corridor import SafeCorridorGenerator
        with pytest.raises(RuntimeError):
            SafeCorridorGenerator(4, 42, 0.3, 4, 4, 1000, 11, 16)


def test_sense_structs_and_argument_checks():
    """hdsm_sense_params layout, hdsm_sense_grid_dims, and hdsm_sense_create rejecting what the kernel cannot do (a ray
    may not cross more than the reference's 1500 voxels; the grid slot must hold the grid)."""
    from multi_agent_pkgs_b200.sensing import HdsmSenseParams, grid_dims
    assert C.sizeof(HdsmSenseParams) == 8 + 24 + 8 + 16
    assert grid_dims(0.3, (20.0, 20.0, 6.0)) == (66, 66, 20)          # floor(range / voxel), map_builder.cpp:103-107
    assert grid_dims(0.2, (6.0, 5.0, 3.0)) == (30, 25, 15)
    L = _lib.load()
    L.hdsm_sense_create.restype = C.c_int
    h = C.c_void_p()

    def create(voxel=0.3, rng=(20.0, 20.0, 6.0), free=0, lim=0, fov=(1.57, 1.57), agents=4, stride=66 * 66 * 20):
        p = HdsmSenseParams(voxel, (C.c_double * 3)(*rng), free, lim, fov[0], fov[1])
        return L.hdsm_sense_create(C.byref(p), C.c_int(agents), C.c_size_t(stride), C.c_int(0), C.byref(h))

    for bad in (dict(voxel=0.0), dict(rng=(20.0, 0.0, 6.0)), dict(rng=(0.1, 20.0, 6.0)), dict(agents=0), dict(stride=1000),
                dict(voxel=0.01, rng=(10.0, 10.0, 10.0), stride=10 ** 9), dict(lim=1, fov=(0.0, 1.0))):
        assert create(**bad) == -1, bad  # HDSM_ERR_INVALID
    assert L.hdsm_sense_create(None, 1, C.c_size_t(8), 0, C.byref(h)) == -1
    if not has_gpu():
        assert create() == -2  # HDSM_ERR_CUDA: no CPU fallback
        from multi_agent_pkgs_b200.sensing import LocalMapBuilder
        with pytest.raises(RuntimeError):
            LocalMapBuilder(0.3, 4)
