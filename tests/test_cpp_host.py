"""The header-only C++ host mirror (xp-game-engine_b200/cpp/xp_physics.hpp) from a native program:
builds on the CPU tier, runs its self-check on the GPU tier (no Python between the program and the C ABI)."""
import os
import shutil
import subprocess

import pytest

from xp_physics_cuda import _lib

CPP = os.path.normpath(os.path.join(_lib.CSRC, "..", "cpp"))


def build_example(tmp_path):
    cxx = shutil.which("g++")
    if cxx is None:
        pytest.skip("no g++")
    _lib.lib()                                   # raises if the product library has not been built
    exe = str(tmp_path / "xp_example")
    r = subprocess.run([cxx, "-std=c++17", "-Wall", "-Werror", "-O1", "-ffp-contract=off", os.path.join(CPP, "example.cpp"), "-L" + _lib.CSRC,
                        "-lxp_physics_cuda", "-Wl,-rpath," + _lib.CSRC, "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_cpp_example_builds_and_links(tmp_path):
    exe = build_example(tmp_path)
    assert os.path.getsize(exe) > 0


@pytest.mark.gpu
def test_cpp_example_self_check(tmp_path):
    exe = build_example(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ok" in r.stdout


# ---- row f1 in C++: the game's simulation step on the library vs on the oracle, natively -----------------
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_controller_parity(tmp_path, oracle):
    cxx = shutil.which("g++")
    if cxx is None:
        pytest.skip("no g++")
    _lib.lib()
    odir = os.path.join(ROOT, "oracle")
    exe = str(tmp_path / "controller_parity")
    r = subprocess.run([cxx, "-std=c++17", "-Wall", "-Werror", "-O1", "-ffp-contract=off",
                        os.path.join(ROOT, "tests", "native", "controller_parity.cpp"),
                        "-L" + _lib.CSRC, "-lxp_physics_cuda", "-L" + odir, "-lxp_oracle",
                        "-Wl,-rpath," + _lib.CSRC, "-Wl,-rpath," + odir, "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_cpp_controller_parity_program_builds(tmp_path, oracle):
    assert os.path.getsize(build_controller_parity(tmp_path, oracle)) > 0


@pytest.mark.gpu
def test_cpp_controller_matches_oracle_frame_by_frame(tmp_path, oracle):
    exe = build_controller_parity(tmp_path, oracle)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ok" in r.stdout
