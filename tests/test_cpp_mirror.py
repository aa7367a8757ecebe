"""The C++ host mirror (quadrotorilqr_b200/cpp/quadrotor_ilqr.hpp) of the reference's C++ interface:
it compiles against include/qilqr.h with plain g++ (CPU check), and the reference's own gtest cases,
ported in tests/cpp/reference_tests.cc, pass when run on the GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "reference_tests.cc")
EXE = os.path.join(ROOT, "tests", "cpp", "reference_tests")


def build_exe():
    from quadrotorilqr_b200 import _capi

    _capi.build()
    lib_dir = os.path.dirname(_capi.LIB_PATH)
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", f'-DQILQR_REPO_ROOT="{ROOT}"', "-o", EXE, SRC, "-L" + lib_dir, "-lqilqr_b200",
           "-Wl,-rpath," + lib_dir, "-L/usr/local/cuda/lib64", "-lcudart"]
    subprocess.check_call(cmd)
    return EXE


def test_cpp_mirror_compiles_and_links():
    assert os.path.exists(build_exe())


@pytest.mark.gpu
def test_reference_gtests_pass_through_the_cpp_mirror():
    exe = build_exe()
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    print(out.stdout)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("[  OK  ]") == 12
