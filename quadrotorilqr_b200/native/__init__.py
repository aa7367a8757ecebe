"""The pybind11 build of the reference's ``quadrotor_ilqr_binding`` module
(``quadrotorilqr_b200/cpp/quadrotor_ilqr_binding.cc``): C++ host code over the C ABI, as the reference's
binding is C++ over ``ILQR<QuadrotorModel>``.

    from quadrotorilqr_b200.native import quadrotor_ilqr_binding
    ilqr = quadrotor_ilqr_binding.QuadrotorILQR(...)      # quadrotor_ilqr_binding.cc:20-49

``build()`` compiles it in-tree with g++ (pybind11 headers from the image; no CUDA code in it).
``quadrotorilqr_b200.quadrotor_ilqr_binding`` is the dependency-free ctypes route to the same library.
"""
from __future__ import annotations

import os
import subprocess
import sysconfig

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "..", "cpp", "quadrotor_ilqr_binding.cc")
MODULE_PATH = os.path.join(_HERE, "quadrotor_ilqr_binding" + sysconfig.get_config_var("EXT_SUFFIX"))


def build(force: bool = False) -> str:
    import pybind11

    from .. import _capi

    _capi.build()
    cpp = os.path.join(_HERE, "..", "cpp")
    deps = [_SRC, os.path.join(cpp, "qilqr_proto.hpp"), os.path.join(cpp, "quadrotor_ilqr.hpp"),
            os.path.join(_HERE, "..", "..", "include", "qilqr.h")]
    stale = force or not os.path.exists(MODULE_PATH) or any(
        os.path.getmtime(d) > os.path.getmtime(MODULE_PATH) for d in deps)
    if stale:
        lib_dir = os.path.dirname(_capi.LIB_PATH)
        subprocess.check_call(
            ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-fvisibility=hidden",
             "-I" + pybind11.get_include(), "-I" + sysconfig.get_paths()["include"], _SRC, "-o", MODULE_PATH,
             "-L" + lib_dir, "-lqilqr_b200", "-Wl,-rpath,$ORIGIN/..", "-L/usr/local/cuda/lib64", "-lcudart"])
    return MODULE_PATH
