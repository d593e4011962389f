"""The ROS2-side shim: include/hdsm_ros_adapter.hpp compiled against stand-ins of the message
classes (ROS2 is not installed here) and the INTEGRATION.md code run as a plain C++ caller of the C ABI."""
import os
import subprocess

import pytest

from conftest import ROOT, has_gpu
from multi_agent_pkgs_b200 import _build

EXE = os.path.join(ROOT, "tests", "ros_adapter_check")


def build_exe():
    lib = _build.build_lib()
    src = os.path.join(ROOT, "tests", "ros_adapter_check.cpp")
    deps = [src, os.path.join(ROOT, "include", "hdsm.h"), os.path.join(ROOT, "include", "hdsm_ros_adapter.hpp")]
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-std=c++14", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), "-o", EXE, src,
                               lib, "-Wl,-rpath," + os.path.dirname(lib)])
    return EXE


def test_adapter_compiles_as_cxx14_and_packs_messages():
    """C++14 like the reference's packages (multi_agent_planner/CMakeLists.txt); message <-> array packing checks."""
    out = subprocess.run([build_exe(), "pack"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    assert "packing ok" in out.stdout
    if not has_gpu():
        assert "hdsm_create returned -2" in out.stdout   # HDSM_ERR_CUDA reaches the shim as a code: no CPU fallback, no crash


@pytest.mark.gpu
def test_shim_plans_two_agents_through_the_c_abi():
    """Two agents, three closed-loop steps with Trajectory 'messages' exchanged in between, from C++ through hdsm_solve_batch."""
    out = subprocess.run([build_exe(), "solve"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "solve ok" in out.stdout
