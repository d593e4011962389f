"""Host-side mirror of the reference's per-update grid post-processing, on top of the C ABI (include/hdsm.h).

Reference: mapping_util/src/map_builder.cpp:207-216 - MapBuilder::SetUncertainToUnknown (:331-365), then
VoxelGrid::InflateObstacles and VoxelGrid::CreatePotentialField (voxel_grid_util/src/voxel_grid.cpp:251-298);
SURVEY.md 8(f) row 4.  `MapProcessor.process` is `hdsm_map_batch`; its output grids are what
`hdsm_corridor_batch` and `hdsm_reftraj_batch` take.

There is no CPU fallback: without the CUDA library / a GPU `MapProcessor` raises.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


class HdsmMapParams(C.Structure):
    _fields_ = [("voxel_size", C.c_double), ("inflation_dist", C.c_double), ("potential_dist", C.c_double),
                ("potential_pow", C.c_int32), ("reserved", C.c_int32)]


class MapProcessor:
    """hdsm_map_create / hdsm_map_batch / hdsm_map_destroy (mapping_util/config defaults: 0.3 / 1.5 / 4)."""

    def __init__(self, voxel_size, max_grids, grid_stride, inflation_dist=0.3, potential_dist=1.5, potential_pow=4, device=0):
        self.L = _lib.load()
        L = self.L
        for f in (L.hdsm_map_create, L.hdsm_map_batch, L.hdsm_map_batch_device):
            f.restype = C.c_int
        L.hdsm_map_last_error.restype = C.c_char_p
        L.hdsm_map_launch_count.restype = C.c_int64
        self.prm = HdsmMapParams(voxel_size, inflation_dist, potential_dist, potential_pow, 0)
        self.grid_stride = int(grid_stride)
        self.h = C.c_void_p()
        rc = L.hdsm_map_create(C.byref(self.prm), C.c_int(max_grids), C.c_size_t(self.grid_stride), C.c_int(device), C.byref(self.h))
        if rc != 0:
            self.h = None
            raise RuntimeError(f"hdsm_map_create failed ({rc}): needs a CUDA device and two copies of a grid in shared memory")

    @property
    def launch_count(self):
        return int(self.L.hdsm_map_launch_count(self.h))

    def process(self, grids):
        """grids [n][dz][dy][dx] int8 -> processed grids of the same shape."""
        grids = np.ascontiguousarray(grids, np.int8)
        n = grids.shape[0]
        assert grids[0].size == self.grid_stride
        dims = np.tile(np.array([grids.shape[3], grids.shape[2], grids.shape[1]], np.int32), (n, 1))
        out = np.empty_like(grids)
        rc = self.L.hdsm_map_batch(self.h, C.c_int(n), grids.ctypes.data_as(C.c_void_p), dims.ctypes.data_as(C.c_void_p),
                                   out.ctypes.data_as(C.c_void_p))
        if rc != 0:
            raise RuntimeError(f"hdsm_map_batch failed ({rc}): {self.L.hdsm_map_last_error(self.h).decode()}")
        return out

    def process_device(self, t_in, t_dims, t_out, stream_ptr=0):
        rc = self.L.hdsm_map_batch_device(self.h, C.c_int(t_in.shape[0]), C.c_void_p(t_in.data_ptr()), C.c_void_p(t_dims.data_ptr()),
                                          C.c_void_p(t_out.data_ptr()), C.c_void_p(stream_ptr))
        if rc != 0:
            raise RuntimeError(f"hdsm_map_batch_device failed ({rc}): {self.L.hdsm_map_last_error(self.h).decode()}")

    def close(self):
        if self.h is not None:
            self.L.hdsm_map_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def raw_local_grid(world, centre, voxel=0.3, grid_range=(20.0, 20.0, 6.0), col_half=0.05):
    """The grid BEFORE post-processing: columns voxelised without inflation, space below z = 0 unknown."""
    from .corridor import local_grid
    return local_grid(world, centre, voxel, grid_range, inflation=0.0, col_half=col_half)
