"""Stand-in headers under oracle/ref_shim/ (test infrastructure that lets parts of the reference compile unmodified here).
The recording stand-in for the Gurobi C++ API is groundwork for compiling agent_class.cpp the way map_builder.cpp already is
(DESIGN.md section 8): this checks that it records what the reference's expression forms mean."""
import os
import subprocess

from conftest import ROOT


def test_recording_gurobi_standin(tmp_path):
    exe = str(tmp_path / "gurobi_standin_check")
    subprocess.check_call(["g++", "-std=c++14", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "oracle", "ref_shim", "gurobi"),
                           "-o", exe, os.path.join(ROOT, "tests", "gurobi_standin_check.cpp")])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stderr
    assert "gurobi stand-in ok" in out.stdout


def test_reference_map_builder_compiles_unmodified_on_the_standins():
    """Where the reference checkout is present: map_builder.cpp + path_tools.cpp + raycast.cpp + voxel_grid.cpp build into
    oracle/_ref/libref_map.so against oracle/ref_shim/ (rclcpp, tf2, message classes, Eigen) without touching a reference file."""
    import pytest
    ref = "/root/reference/mapping_util/src/map_builder.cpp"
    if not os.path.exists(ref):
        pytest.skip("reference checkout not present")
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "-B", "_ref/libref_map.so"])
    from oracle import sensing as S
    assert S.have_ref()
