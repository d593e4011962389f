"""Builds libhdsm.so in-tree with nvcc for sm_100a (no torch involved)."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libhdsm.so")
SOURCES = ["hdsm_capi.cu", "hdsm_corridor.cu", "hdsm_reftraj.cu", "hdsm_map.cu", "hdsm_sense.cu"]
DEPS = ["hdsm_capi.cu", "hdsm_corridor.cu", "hdsm_reftraj.cu", "hdsm_map.cu", "hdsm_sense.cu", "hdsm_sense_core.h", "hdsm_common.h", "hdsm_kernel.cuh", "hdsm_tables.h", os.path.join("..", "..", "include", "hdsm.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-diag-suppress", "68", "-shared", "-Xcompiler", "-fPIC,-pthread"]


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build_lib(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/ into multi_agent_pkgs_b200/libhdsm.so if it is missing or older than its sources."""
    if force or stale():
        cmd = [_nvcc(), *NVCC_FLAGS, "-o", LIB, *[os.path.join(CSRC, s) for s in SOURCES], "-ldl"]
        if os.environ.get("HDSM_MINBLOCKS"):
            cmd.insert(1, "-DHDSM_MINBLOCKS=" + os.environ["HDSM_MINBLOCKS"])
        if os.environ.get("HDSM_ENABLE_PROFILE"):
            cmd.insert(1, "-DHDSM_ENABLE_PROFILE")
        if verbose:
            cmd.insert(1, "-Xptxas")
            cmd.insert(2, "-v")
        subprocess.check_call(cmd)
    return LIB
